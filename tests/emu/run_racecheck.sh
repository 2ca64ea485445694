#!/bin/bash
# TEST-ONLY: schedule-permutation race check of the product kernels without a GPU.  The host build of the kernels resumes
# the runnable threads of a block in reversed and in randomly shuffled order (TNL_EMU_SCHED, tests/emu/cuda_shim.h); a
# kernel that is missing a __syncthreads / __syncwarp between a shared-memory write and a read by another thread changes
# its result with the order and fails its oracle comparison.
#   tests/emu/run_racecheck.sh [pytest args]
set -e
cd "$(dirname "$0")/../.."
python tests/emu/gen_kemu.py
if [ $# -eq 0 ]; then set -- tests/test_kernels_emu.py tests/test_host_on_emu.py -k "not world2"; fi
for sched in reverse random:1 random:2; do
    echo "== TNL_EMU_SCHED=$sched"
    TNL_EMU_SCHED=$sched python -m pytest -x -q -p no:cacheprovider "$@"
done
