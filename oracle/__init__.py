"""TEST INFRASTRUCTURE -- CPU oracle for the finite-volume hot path (see oracle/fvo.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. The product (foamadapter_b200) never does.
"""
