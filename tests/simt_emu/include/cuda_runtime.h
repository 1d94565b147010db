// shim: `#include <cuda_runtime.h>` of the engine sources resolves here in the host-side SIMT emulation build (tests only)
#pragma once
#include "simt_emu.h"
