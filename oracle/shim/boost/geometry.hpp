#pragma once
#include <algorithm>
#include "shim_boost.hpp"
