// Shim for <Random123/threefry.h>. RandBLAS includes it (random_gen.hh:38) but the
// sketch-and-factor path only instantiates Philox4x32 (base.hh:53); intentionally empty.
#pragma once
#include "array.h"
