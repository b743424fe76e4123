// Same-named stand-in for the reference's dev/attention.cuh: with include/legacy in front of the include path a reference
// translation unit compiles against libunet_b200.so without a source change (declarations: ../unet_b200_legacy.hpp).
#pragma once
#include "../unet_b200_legacy.hpp"
