"""TEST-ONLY: builds tests/emu/_build/libkemu.so -- the PRODUCT kernels of trinerflet_b200/csrc/{raymarch,sample,sort,tiles,
optim,grid,rays}.cu compiled for the host on top of tests/emu/cuda_shim.h (fibers for the threads of a block, real
__syncthreads / warp-collective rendezvous), exporting the same `tnl_*` C ABI as libtrinerflet_b200.so but taking HOST
pointers.  The CPU test-suite calls it to check the kernels' code against the oracle without a GPU.

The product sources are used as they are; this script rewrites only what a host compiler cannot parse:
  * every launch   kernel<targs><<<grid, block, smem, stream>>>(args);
        ->         tnl_emu::launch(tnl_emu::cfg(grid, block, smem, stream), [&]() { kernel<targs>(args); });
  * the one inline-PTX statement of sample.cu (`red.global.add.v4.f32`) -> four float additions;
  * `extern __shared__ T name[];` -> `T* name = (T*)tnl_emu::dyn_smem();` (sized by the launch's shared-memory argument).
Nothing here is imported by the product; the library is never shipped as part of it.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "trinerflet_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
SOURCES = ["api.cu", "raymarch.cu", "sample.cu", "sort.cu", "tiles.cu", "optim.cu", "grid.cu", "rays.cu", "idwt.cu", "mlp.cu"]
# TNL_KEMU_ASAN=1: AddressSanitizer build (run the tests with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
# ASAN_OPTIONS=detect_leaks=0): out-of-bounds reads / writes of the kernels on the callers' heap buffers become hard errors
ASAN = os.environ.get("TNL_KEMU_ASAN") == "1"
SO = os.path.join(OUT_DIR, "libkemu_asan.so" if ASAN else "libkemu.so")


def _match_back_template(s, end):
    """s[end-1] == '>' : index of the matching '<' (template argument list of the kernel name)"""
    depth = 0
    i = end - 1
    while i >= 0:
        if s[i] == '>':
            depth += 1
        elif s[i] == '<':
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template argument list before <<<")


def _match_paren(s, start):
    """s[start] == '(' : index just past the matching ')'"""
    depth = 0
    i = start
    while i < len(s):
        if s[i] == '(':
            depth += 1
        elif s[i] == ')':
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced parentheses after >>>")


def rewrite_launches(src):
    out = []
    pos = 0
    n = 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            break
        # kernel name expression: identifier (with ::) optionally followed by <template args>
        e = k
        while e > 0 and src[e - 1].isspace():
            e -= 1
        b = e
        if src[b - 1] == '>':
            b = _match_back_template(src, b)
        while b > 0 and (src[b - 1].isalnum() or src[b - 1] in "_:"):
            b -= 1
        name = src[b:e]
        c_end = src.find(">>>", k)
        cfg = src[k + 3:c_end]
        a0 = c_end + 3
        while src[a0].isspace():
            a0 += 1
        assert src[a0] == '(', (name, src[a0:a0 + 20])
        a1 = _match_paren(src, a0)
        args = src[a0:a1]
        out.append(src[pos:b])
        out.append(f"tnl_emu::launch(tnl_emu::cfg({cfg}), [&]() {{ {name}{args}; }})")
        pos = a1
        n += 1
    return "".join(out), n


_RED = re.compile(r'asm volatile\("red\.global\.add\.v4\.f32[^;]*?;\\n"\s*::\s*"l"\((?P<addr>\w+)\),\s*"f"\((?P<a>[^"]+?)\),\s*"f"\((?P<b>[^"]+?)\),'
                  r'\s*"f"\((?P<c>[^"]+?)\),\s*"f"\((?P<d>[^"]+?)\)\s*:\s*"memory"\);', re.S)


# the three warp-wide instructions of the mma.sync MLP kernels (csrc/mlp.cu); the helper functions that wrap them name their
# operands d / a / b and r / p
_MMA = re.compile(r'asm volatile\(\s*"mma\.sync\.aligned\.m16n8k16\.row\.col\.f32\.f16\.f16\.f32[^;]*?;\\n"[^;]*?\);', re.S)
_LDSM = re.compile(r'asm volatile\("ldmatrix\.sync\.aligned\.m8n8\.x(?P<n>[24])\.trans\.shared\.b16[^;]*?;\\n"[^;]*?\);', re.S)


def rewrite_ptx(src):
    def sub(m):
        g = m.groupdict()
        return (f"{{ float* a_ = {g['addr']}; a_[0] += {g['a']}; a_[1] += {g['b']}; a_[2] += {g['c']}; a_[3] += {g['d']}; }}")
    src, n = _RED.subn(sub, src)
    src, k = _MMA.subn("tnl_emu::mma_m16n8k16(d, a, b.x, b.y);", src)
    n += k
    src, k = _LDSM.subn(lambda m: f"tnl_emu::ldmatrix_trans<{m.group('n')}>(r, p);", src)
    n += k
    # device-only instruction wrappers that come with a host definition: `#ifdef __CUDACC__ <asm> #else <host> #endif`
    src, k = re.subn(r'#ifdef __CUDACC__\n(?:(?!#else|#endif).)*?asm volatile.*?#else[^\n]*\n(.*?)#endif', r'\1', src, flags=re.S)
    n += k
    if "asm" in re.sub(r"//.*", "", src):
        raise RuntimeError("inline PTX the emulator does not know")
    return src, n


_DYN_SMEM = re.compile(r'extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];')


# the tcgen05 implementation (csrc/mlp_tc.cu) cannot be emulated: the host build reports it as unsupported, so the C ABI
# dispatches every call to the mma.sync kernels of csrc/mlp.cu
MLP_TC_STUB = '''// GENERATED by tests/emu/gen_kemu.py -- test-only
#include "cuda_shim.h"
#include "mlp_tc.cuh"
#include <stdlib.h>
namespace tnl {
bool mlp_tc_supported(uint32_t, uint32_t, uint32_t) { return false; }
size_t mlp_tc_packed_bytes(uint32_t, uint32_t) { return 0; }
void mlp_tc_pack(uint32_t, uint32_t, const float*, const float*, const float*, const float*, const float*, void*, cudaStream_t) { abort(); }
void mlp_tc_forward(uint32_t, uint32_t, const void*, const void*, const float*, uint32_t, const int32_t*, float*, float*, float*, cudaStream_t) { abort(); }
void mlp_tc_backward(uint32_t, uint32_t, const void*, const void*, const float*, uint32_t, const int32_t*, const float*, const float*, void*,
                     float*, float*, float*, float*, float*, cudaStream_t) { abort(); }
}  // namespace tnl
'''


def generate(name):
    src = open(os.path.join(CSRC, name)).read()
    src = _DYN_SMEM.sub(r'\1* \2 = reinterpret_cast<\1*>(tnl_emu::dyn_smem());', src)
    src, n_launch = rewrite_launches(src)
    src, n_ptx = rewrite_ptx(src)
    dst = os.path.join(OUT_DIR, ("kemu_asan_" if ASAN else "kemu_") + name.replace(".cu", ".cpp"))
    with open(dst, "w") as f:
        f.write(f"// GENERATED by tests/emu/gen_kemu.py from trinerflet_b200/csrc/{name} ({n_launch} launches, {n_ptx} PTX statements "
                f"rewritten) -- test-only\n#include \"cuda_shim.h\"\n#line 1 \"{os.path.join(CSRC, name)}\"\n")
        f.write(src)
    return dst


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, h) for h in ("common.cuh", "rays_core.cuh", "idwt_core.cuh", "mlp_math.cuh", "mlp_tc.cuh", "sample_coords.cuh")]
    deps += [os.path.join(HERE, "cuda_shim.h"), os.path.abspath(__file__), os.path.join(ROOT, "include", "trinerflet_b200.h")]
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(d) for d in deps):
        return SO
    objs = []
    procs = []
    for s in SOURCES:
        cpp = generate(s)
        obj = cpp.replace(".cpp", ".o")
        objs.append(obj)
        san = ["-fsanitize=address", "-fno-omit-frame-pointer"] if ASAN else []
        procs.append(subprocess.Popen(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", *san, "-I", HERE,
                                       "-I", os.path.join(HERE, "shim_include"), "-I", CSRC, "-c", cpp, "-o", obj]))
    stub = os.path.join(OUT_DIR, ("kemu_asan_" if ASAN else "kemu_") + "mlp_tc_stub.cpp")
    with open(stub, "w") as f:
        f.write(MLP_TC_STUB)
    objs.append(stub.replace(".cpp", ".o"))
    procs.append(subprocess.Popen(["g++", "-O1", "-std=c++17", "-fPIC", "-w", *(["-fsanitize=address"] if ASAN else []), "-I", HERE,
                                   "-I", os.path.join(HERE, "shim_include"), "-I", CSRC, "-c", stub, "-o", objs[-1]]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("kernel emulator: compilation failed")
    subprocess.check_call(["g++", "-shared", *(["-fsanitize=address"] if ASAN else []), "-o", SO] + objs)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv)
