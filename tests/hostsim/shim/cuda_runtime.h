// TEST INFRASTRUCTURE: stands in for <cuda_runtime.h> when the product sources are compiled for the CPU (tests/hostsim).
#pragma once
#include "../cuda_sim.hpp"
