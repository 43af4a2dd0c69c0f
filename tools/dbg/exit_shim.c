#define _GNU_SOURCE
#include <execinfo.h>
#include <stdio.h>
#include <unistd.h>
#include <stdlib.h>
void exit(int c)
{
    void* b[64];
    int n = backtrace(b, 64);
    fprintf(stderr, "exit(%d) called from:\n", c);
    backtrace_symbols_fd(b, n, 2);
    _exit(c);
}
