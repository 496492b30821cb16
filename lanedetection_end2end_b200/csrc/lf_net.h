// Internal aliases for the public argument structs.
#pragma once
#include "../../include/lanefit_b200.h"
namespace lf {
using ConvArgs = LfConvArgs;
using WgradArgs = LfWgradArgs;
}  // namespace lf
