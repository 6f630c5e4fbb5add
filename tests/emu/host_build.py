"""Build host (CPU) libraries / executables out of the repo's ``.cu`` files for the emulation tests
(tests/test_kernel_emulation.py, test_attention_kernel_model.py, test_gemm_kernel_model.py; overview in
docs/guide/testing.md).

The kernel source is used as is.  The textual changes are:
  * the launch syntax, which a C++ compiler cannot parse: ``kernel<<<grid, threads, smem, stream>>>(args)`` becomes
    ``cuda_emu::launch[_dyn | _cluster2](dim3(grid), dim3(threads).x, [smem,] [&] { kernel(args); })``, so the real
    ``extern "C"`` launchers (grid / block selection, dtype and flag dispatch, tensor-map construction) run too;
  * the handful of inline-PTX statements listed in ``_PTX`` below, replaced by their meaning on the model;
  * ``extern __shared__`` (dynamic shared memory) becomes the block's buffer.
Everything else -- threadIdx, __shared__, __syncthreads, shuffles, the 16-bit types (cuda_emu/cuda_emu.h); mbarrier,
TMA, tensor memory, tcgen05.mma, clusters (cuda_emu/tcgen05_model.h, ptx.cuh) -- comes from headers that shadow the
CUDA ones on the include path.  The descriptor builders of the real ptx.cuh are copied verbatim (``ptx_real_extract.h``)
so the model decodes exactly what the kernels encode."""
import os
import re
import subprocess

EMU = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(EMU)), "megatron_llm_b200", "csrc")
_LAUNCH = re.compile(r"((?:mlb::)?\w+(?:<[^<>;()]*>)?)\s*<<<")
# the inline-PTX statements of the translated files, replaced by their host / model meaning
_PTX = [(re.compile(r'asm volatile\("red\.global\.add\.v4\.f32 \[%0\], \{%1, %2, %3, %4\};"\s*::\s*"l"\((\w+)\), "f"\((\w+)\), '
                    r'"f"\((\w+)\), "f"\((\w+)\), "f"\((\w+)\)\s*:\s*"memory"\);'),
         r"cuda_emu::atomic_add4(\1, \2, \3, \4, \5);"),
        # NVLS (csrc/comm.cu): the in-switch reduction / multicast store, emulated over the registered copies
        (re.compile(r'asm volatile\("multimem\.ld_reduce\.relaxed\.sys\.global\.add\.v4\.f32 \{%0, %1, %2, %3\}, \[%4\];"\s*'
                    r':\s*"=f"\((v\[u\])\.x\), "=f"\(v\[u\]\.y\), "=f"\(v\[u\]\.z\), "=f"\(v\[u\]\.w\)\s*'
                    r':\s*"l"\(([^;]*?)\)\s*:\s*"memory"\);', re.S),
         r"\1 = cuda_emu::multimem_ld_reduce_add_v4(\2);"),
        (re.compile(r'asm volatile\("multimem\.st\.relaxed\.sys\.global\.v4\.f32 \[%0\], \{%1, %2, %3, %4\};"\s*::\s*'
                    r'"l"\(([^;]*?)\),\s*"f"\(([\w.]+)\), "f"\(([\w.]+)\), "f"\(([\w.]+)\), "f"\(([\w.]+)\)\s*'
                    r':\s*"memory"\);', re.S),
         r"cuda_emu::multimem_st_v4(\1, \2, \3, \4, \5);"),
        # attention kernels (csrc/attention_common.cuh, attention_bwd_sm100.cu): exp2, the 16-column tcgen05.st, the
        # named barrier of a 128-thread warpgroup -- on the functional model of tcgen05_model.h
        (re.compile(r'asm volatile\("ex2\.approx\.ftz\.f32 %0, %1;" : "=f"\((\w+)\) : "f"\((\w+)\)\);'), r"\1 = exp2f(\2);"),
        (re.compile(r'asm volatile\(\s*"tcgen05\.st\.sync\.aligned\.32x32b\.x16\.b32.*?:\s*"memory"\);', re.S),
         r"mlb::tmem_st_n<16>(taddr, r);"),
        (re.compile(r'asm volatile\("bar\.sync (\d+), (\d+);" ::: "memory"\);'), r"cuda_emu::named_barrier(\1, \2);"),
        (re.compile(r'asm volatile\("mov\.u64 %0, %%clock64;" : "=l"\((\w+)\)\);'), r"\1 = 0;"),
        (re.compile(r"extern __shared__ uint8_t smem_raw\[\];"), r"uint8_t* smem_raw = cuda_emu::bm->smem;"),
        # GEMM header (csrc/gemm_sm100.cuh): the bulk copies / multicast store of the fused modes.  Those modes need
        # co-resident CTAs and peers and are NOT run on the model; the statements only have to compile.
        (re.compile(r'asm volatile\("cp\.async\.bulk\.shared::cluster\.global\.mbarrier::complete_tx::bytes.*?:\s*"memory"\);', re.S),
         r"std::memcpy(smem_dst, gsrc, bytes); mbar_complete_tx(bar, bytes);"),
        (re.compile(r'asm volatile\("cp\.async\.bulk\.global\.shared::cta\.bulk_group.*?:\s*"memory"\);', re.S),
         r"std::memcpy(gdst, smem_src, bytes);"),
        (re.compile(r'asm volatile\("multimem\.st\.relaxed\.sys\.global\.v4\.f32 \[%0\], \{%1, %2, %3, %4\};" ::"l"\(mc_addr\),.*?:\s*"memory"\);', re.S),
         r"cuda_emu::multimem_st_v4(static_cast<float*>(mc_addr), __uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));")]


def _split_top_level(text):
    parts, depth, cur = [], 0, []
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


def launches_to_host(src: str) -> str:
    for pat, repl in _PTX:
        src = pat.sub(repl, src)
    cluster2 = "__cluster_dims__(2, 1, 1)" in src          # every kernel of such a file is launched as CTA pairs
    out, i = [], 0
    while True:
        m = _LAUNCH.search(src, i)
        if not m:
            out.append(src[i:])
            return "".join(out)
        out.append(src[i:m.start()])
        j = src.index(">>>", m.end())
        cfg = _split_top_level(src[m.end():j])
        k = src.index("(", j)
        depth = 0
        for e in range(k, len(src)):
            depth += src[e] == "("
            depth -= src[e] == ")"
            if depth == 0:
                break
        call = f"[&] {{ {m.group(1)}({src[k + 1:e]}); }}"
        if cluster2:
            out.append(f"cuda_emu::launch_cluster2(dim3({cfg[0]}), dim3({cfg[1]}).x, {cfg[2]}, {call})")
        elif len(cfg) > 2 and cfg[2] != "0":      # dynamic shared memory: a kernel of the tcgen05 / TMA model
            out.append(f"cuda_emu::launch_dyn(dim3({cfg[0]}), dim3({cfg[1]}).x, {cfg[2]}, {call})")
        else:
            out.append(f"cuda_emu::launch(dim3({cfg[0]}), dim3({cfg[1]}).x, {call})")
        i = e + 1


_MODEL_HEADERS = ("attention_common.cuh", "gemm_sm100.cuh")       # csrc headers with inline PTX: transformed copies shadow the originals


_SHARED_SCALAR = re.compile(r"__shared__\s+(int|unsigned|uint32_t|float|bool)\s+(\w+)\s*;")


def _per_block_shared_scalars(src):
    """`__shared__ int x;` -> one instance per block (the files of the tcgen05 model run CTA pairs / all blocks of a
    launch concurrently, where the `static` of the plain __shared__ macro would be shared between CTAs)."""
    return _SHARED_SCALAR.sub(lambda m: f"{m.group(1)}& {m.group(2)} = cuda_emu::block_static<{m.group(1)}>(__LINE__);", src)


def _write_model_headers(out_dir):
    """For the kernels that run on the functional tcgen05 / TMA model: a transformed copy of the csrc headers that carry
    inline PTX, and the descriptor builders of the REAL ptx.cuh (the model decodes what the real code encodes)."""
    for h in _MODEL_HEADERS:
        with open(os.path.join(out_dir, h), "w") as fh:
            fh.write(_per_block_shared_scalars(launches_to_host(open(os.path.join(CSRC, h)).read())))
    real = open(os.path.join(CSRC, "ptx.cuh")).read()
    a = real.index("// ---------------------------------------------------------------- UMMA descriptors")
    b = real.index("// ---------------------------------------------------------------- cluster")
    with open(os.path.join(out_dir, "ptx_real_extract.h"), "w") as fh:
        fh.write("// extracted verbatim from csrc/ptx.cuh by tests/emu/host_build.py\n#pragma once\nnamespace mlb {\n"
                 + real[a:b] + "}  // namespace mlb\n")


def _host_sources(cu_files, out_dir, mutate=None):
    os.makedirs(out_dir, exist_ok=True)
    _write_model_headers(out_dir)
    sources = []
    for f in cu_files:
        body = open(os.path.join(CSRC, f)).read()
        if mutate is not None:
            body = mutate(f, body)
        body = launches_to_host(body)
        assert "<<<" not in body and "asm volatile" not in body, f
        dst = os.path.join(out_dir, f.replace(".cu", "_host.cpp"))
        with open(dst, "w") as fh:
            fh.write('#include "cuda_emu.h"\n' + body)
        sources.append(dst)
    return sources


def build(cu_files, out_dir, name="emu_kernels", extra_cpp=()):
    """Compile ``cu_files`` (names inside csrc/) plus ``extra_cpp`` (paths) into ``<out_dir>/<name>.so``."""
    sources = _host_sources(cu_files, out_dir)
    so = os.path.join(out_dir, name + ".so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-shared", "-fPIC", "-pthread", "-w",
                           "-I" + out_dir, "-I" + os.path.join(EMU, "cuda_emu"), "-I" + CSRC, *sources, *extra_cpp, "-o", so])
    return so


def build_race_driver(cu_files, out_dir, name="race_driver", mutate=None, defines=(), driver="race_driver.cpp"):
    """``race_driver.cpp`` + the kernels as an executable instrumented by ThreadSanitizer (see the driver's header).
    ``mutate(file_name, source) -> source`` lets a test break a kernel on purpose to show that the race is found."""
    sources = _host_sources(cu_files, out_dir, mutate)
    exe = os.path.join(out_dir, name)
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++20", "-fsanitize=thread", "-pthread", "-w",
                           *["-D" + d for d in defines], "-I" + out_dir, "-I" + os.path.join(EMU, "cuda_emu"), "-I" + CSRC, *sources,
                           os.path.join(EMU, driver), "-o", exe])
    return exe


def build_executable(cu_files, driver, out_dir, name):
    """The kernels plus a C++ driver from tests/emu as a plain executable (one emulated rank per process)."""
    sources = _host_sources(cu_files, out_dir)
    exe = os.path.join(out_dir, name)
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-w", "-I" + out_dir,
                           "-I" + os.path.join(EMU, "cuda_emu"), "-I" + CSRC, *sources, os.path.join(EMU, driver), "-o", exe])
    return exe
