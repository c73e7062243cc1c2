#pragma once
#include "../../include/avocado_b200_synth.h"
