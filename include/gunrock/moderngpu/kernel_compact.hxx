// moderngpu/kernel_compact.hxx -- forwards to the B200 engine-backed subset of the mgpu host API.
#pragma once
#include "../mgpu_compat.hxx"
