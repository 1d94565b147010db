#!/bin/bash
# Host-side sanitizers over the kernel sources (no GPU needed): builds the SIMT-emulated library (tests/simt_emu) with
# UBSan (incl. -fsanitize=bounds-strict) and with ASan, and runs tests/test_simt_emu.py against each.  Device buffers are
# heap allocations here (cudaMalloc -> posix_memalign, numpy arrays), so ASan's red zones catch out-of-bounds global-memory
# and dynamic-shared-memory accesses of any kernel the tests launch; UBSan catches misaligned vector accesses, bad shifts,
# signed overflow and static-array overruns.  compute-sanitizer on the B200 (tools/sanitize.py) remains the check for
# races and the real memory model.
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import os, subprocess, sys
sys.path.insert(0, "tests")
from simt_emu import build as b
b.build()
base = [f for f in b.CXXFLAGS if f != "-O2"] + ["-O1", "-I", os.path.join(b.HERE, "include")]
src = os.path.join(b.BUILD, "src", "ses_abi.cpp")
subprocess.check_call(["g++"] + base + ["-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-fsanitize=bounds-strict",
                                        "-o", os.path.join(b.BUILD, "libses_simt_emu_ubsan.so"), src])
subprocess.check_call(["g++"] + base + ["-fsanitize=address", "-fno-omit-frame-pointer",
                                        "-o", os.path.join(b.BUILD, "libses_simt_emu_asan.so"), src])
PY
B=$PWD/tests/simt_emu/_build
echo "== UBSan"; SES_SIMT_EMU_LIB=$B/libses_simt_emu_ubsan.so UBSAN_OPTIONS=print_stacktrace=1 python -m pytest tests/test_simt_emu.py tests/test_gru_generic.py -m "not gpu" -x -q -p no:cacheprovider | tail -2
echo "== ASan";  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    SES_SIMT_EMU_LIB=$B/libses_simt_emu_asan.so python -m pytest tests/test_simt_emu.py tests/test_gru_generic.py -m "not gpu" -x -q -p no:cacheprovider | tail -2
