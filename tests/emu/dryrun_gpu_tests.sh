#!/bin/bash
# TEST-ONLY developer tool: dry-run GPU test files on a machine without a GPU.  Copies the given tests/test_gpu_*.py files to
# a scratch directory with the device strings mapped to the CPU and runs them on CPU tensors over the host build of the
# kernels (tests/emu_backend.py).  It checks the TEST LOGIC and the host code (shapes, argument order, assertions) before GPU
# time is spent on them; it proves nothing about the device build.  Expected artefacts: comparisons against the compiled
# reference CUDA extensions (oracle/_ref) and anything that needs fp16 autocast / the tcgen05 kernels fail or are skipped.
#   tests/emu/dryrun_gpu_tests.sh tests/test_gpu_x_feeder.py tests/test_gpu_x_infer_loop.py [-- pytest args]
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
OUT="$(mktemp -d /tmp/tnl_dryrun.XXXXXX)"
files=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do files+=("$1"); shift; done
[ "$1" == "--" ] && shift
for f in "${files[@]}"; do
    sed -e 's/device="cuda"/device="cpu"/g' -e 's/dev = "cuda"/dev = "cpu"/g' -e 's/torch.device("cuda")/torch.device("cpu")/g' \
        -e 's/, "cuda")/, "cpu")/g' -e 's/pytestmark = pytest.mark.gpu/pytestmark = []/' \
        -e 's/torch.Generator(device="cpu")/torch.Generator()/g' -e 's/torch.Generator(device=dev)/torch.Generator()/g' \
        "$ROOT/$f" > "$OUT/$(basename "${f%.py}")_cpu.py"
done
cat > "$OUT/conftest.py" <<PY
import sys
sys.path.insert(0, "$ROOT")
import pytest, torch
from tests import emu_backend

@pytest.fixture(autouse=True)
def _emu(monkeypatch):
    emu_backend.install(monkeypatch)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    yield

@pytest.fixture(scope="session")
def golden_dir():
    return "$ROOT/tests/golden"
PY
cd "$OUT" && python -m pytest -q -p no:cacheprovider "$@" .
