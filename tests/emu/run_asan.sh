#!/bin/bash
# TEST-ONLY: memcheck of the product kernels without a GPU.  Builds the host build of the kernels with AddressSanitizer
# (tests/emu/gen_kemu.py, TNL_KEMU_ASAN=1) and runs the emulator-based CPU tests under it: an out-of-bounds read or write of a
# kernel on the caller's buffers aborts with the kernel's source line (the generated files carry #line directives).
#   tests/emu/run_asan.sh [pytest args]        default: tests/test_kernels_emu.py tests/test_host_on_emu.py -k "not world2"
# Tests that expect a C++ exception out of torch (pytest.raises over a torch error, the double-backward check of the position
# gradient) abort under a preloaded libasan (its __cxa_throw interceptor is unresolved in a non-ASAN python): deselect them, e.g.
#   tests/emu/run_asan.sh tests/test_sr_encoder.py -k "not position and not constructor"
set -e
cd "$(dirname "$0")/../.."
export TNL_KEMU_ASAN=1
python tests/emu/gen_kemu.py
if [ $# -eq 0 ]; then set -- tests/test_kernels_emu.py tests/test_host_on_emu.py -k "not world2"; fi
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    python -m pytest -x -q -p no:cacheprovider "$@"
