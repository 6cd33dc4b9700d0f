#pragma once
#include "shim_ompl.hpp"
