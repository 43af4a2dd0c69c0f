// Umbrella header of the B200 build of the NeoN API surface (reference: src/NeoN/include/NeoN/NeoN.hpp).
#pragma once
#include "NeoN/core.hpp"
#include "NeoN/mesh.hpp"
#include "NeoN/finiteVolume.hpp"
#include "NeoN/linearAlgebra.hpp"
#include "NeoN/dsl.hpp"
