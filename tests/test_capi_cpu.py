"""CPU: the C-ABI library loads and exports every symbol the headers under include/ declare; struct layouts used by the
Python binding match the C definitions (sizes probed by compiling a tiny C program against the headers); error paths
return CUVS_ERROR with text and never throw.  No GPU compute calls."""
import ctypes as C
import glob
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(INC, "**", "*.h"), recursive=True):
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"CUVS_EXPORT\s+[\w\s\*]+?\b(cuvs\w+)\s*\(", txt):
            syms.add(m.group(1))
    return sorted(syms)


@pytest.fixture(scope="module")
def lib():
    from cuvs_b200 import build
    path = build.build()
    return C.CDLL(path)


def test_every_declared_symbol_is_exported(lib):
    syms = _declared_symbols()
    assert len(syms) > 100
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_headers_compile_as_plain_c_and_struct_sizes_match_binding():
    """c/tests/core/headers.c analogue: the headers are pure C; and the ctypes mirrors have the compiler's sizes."""
    src = r'''
#include <stdio.h>
#include <cuvs/core/all.h>
#include <cuvs_b200/ext.h>
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(struct cuvsIvfFlatIndexParams), sizeof(struct cuvsIvfPqIndexParams),
         sizeof(struct cuvsIvfPqSearchParams), sizeof(struct cuvsCagraIndexParams), sizeof(struct cuvsCagraSearchParams),
         sizeof(struct cuvsKMeansParams), sizeof(DLManagedTensor), sizeof(cuvsBruteForceIndex), sizeof(cuvsFilter));
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "h.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "h")
        subprocess.run(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-I", INC, "-I", "/usr/local/cuda/include", c, "-o", exe], check=True)
        sizes = list(map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()))
    from cuvs_b200 import _capi
    from cuvs_b200.cluster import kmeans
    from cuvs_b200.neighbors import cagra, ivf_flat, ivf_pq
    ours = [C.sizeof(ivf_flat._IndexParamsC), C.sizeof(ivf_pq._IndexParamsC), C.sizeof(ivf_pq._SearchParamsC), C.sizeof(cagra._IndexParamsC),
            C.sizeof(cagra._SearchParamsC), C.sizeof(kmeans._ParamsC), C.sizeof(_capi.DLManagedTensor), C.sizeof(_capi.index_handle),
            C.sizeof(_capi.cuvsFilter)]
    assert ours == sizes


def test_param_defaults_match_reference(lib):
    """c/src/neighbors/{ivf_flat,ivf_pq,cagra}.cpp ParamsCreate defaults."""
    from cuvs_b200.neighbors import cagra, ivf_flat, ivf_pq
    p = C.POINTER(ivf_flat._IndexParamsC)()
    assert lib.cuvsIvfFlatIndexParamsCreate(C.byref(p)) == 1
    c = p.contents
    assert (c.metric, c.metric_arg, c.add_data_on_build, c.n_lists, c.kmeans_n_iters, c.kmeans_trainset_fraction) == (0, 2.0, True, 1024, 20, 0.5)
    lib.cuvsIvfFlatIndexParamsDestroy(p)
    p = C.POINTER(ivf_pq._IndexParamsC)()
    assert lib.cuvsIvfPqIndexParamsCreate(C.byref(p)) == 1
    c = p.contents
    assert (c.pq_bits, c.pq_dim, c.codebook_kind, c.max_train_points_per_pq_code, c.codes_layout) == (8, 0, 0, 256, 1)
    lib.cuvsIvfPqIndexParamsDestroy(p)
    p = C.POINTER(ivf_pq._SearchParamsC)()
    assert lib.cuvsIvfPqSearchParamsCreate(C.byref(p)) == 1
    c = p.contents
    assert (c.n_probes, c.lut_dtype, c.internal_distance_dtype, c.max_internal_batch_size, c.preferred_shmem_carveout) == (20, 0, 0, 4096, 1.0)
    lib.cuvsIvfPqSearchParamsDestroy(p)
    p = C.POINTER(cagra._SearchParamsC)()
    assert lib.cuvsCagraSearchParamsCreate(C.byref(p)) == 1
    c = p.contents
    assert (c.itopk_size, c.search_width, c.num_random_samplings, c.rand_xor_mask) == (64, 1, 1, 0x128394)
    assert abs(c.hashmap_max_fill_rate - 0.5) < 1e-7
    lib.cuvsCagraSearchParamsDestroy(p)


def test_errors_are_codes_with_text_not_exceptions(lib):
    lib.cuvsGetLastErrorText.restype = C.c_char_p
    h = C.c_size_t(0)
    rc = lib.cuvsResourcesCreate(C.byref(h))  # no GPU in this container -> CUVS_ERROR + message
    if rc == 0:
        assert b"CUDA" in lib.cuvsGetLastErrorText()
    lib.cuvsSetLastErrorText(None)
    assert lib.cuvsGetLastErrorText() is None
    lib.cuvsSetLastErrorText(b"boom")
    assert lib.cuvsGetLastErrorText() == b"boom"
    major, minor, patch = C.c_uint16(), C.c_uint16(), C.c_uint16()
    assert lib.cuvsVersionGet(C.byref(major), C.byref(minor), C.byref(patch)) == 1 and (major.value, minor.value) == (26, 8)
    lib.cuvsSetLogLevel(4)
    assert lib.cuvsGetLogLevel() == 4


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under cuvs_b200/ may reference it."""
    bad = []
    for f in glob.glob(os.path.join(ROOT, "cuvs_b200", "**", "*"), recursive=True):
        if os.path.isfile(f) and f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp")):
            txt = open(f, errors="ignore").read()
            if re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M) or "liboracle" in txt:
                bad.append(f)
    assert not bad, bad


def test_plain_c_client_compiles_links_and_runs_without_a_gpu(tmp_path):
    """examples/c/ivf_pq_search.c uses only the reference's C API (what a cgo / JNI / bindgen binding calls): it must
    compile as C against include/, link against the in-tree libcuvs_c.so and run its GPU-free path."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "ivf_pq_search"
    cmd = ["/usr/bin/gcc", os.path.join(root, "examples", "c", "ivf_pq_search.c"), "-I" + os.path.join(root, "include"),
           "-I/usr/local/cuda/include", "-L" + os.path.join(root, "cuvs_b200", "lib"), "-lcuvs_c", "-L/usr/local/cuda/lib64",
           "-lcudart", "-Wl,-rpath," + os.path.join(root, "cuvs_b200", "lib"), "-Wall", "-Werror", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe), "--version"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("libcuvs_c 26."), out.stdout + out.stderr


def _build_cpp_client(out):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = ["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", os.path.join(root, "examples", "cpp", "ivf_pq_search.cpp"),
           "-I" + os.path.join(root, "include"), "-I/usr/local/cuda/include", "-L" + os.path.join(root, "cuvs_b200", "lib"), "-lcuvs_c",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.join(root, "cuvs_b200", "lib"), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_header_adaptor_compiles_links_and_runs_without_a_gpu(tmp_path):
    """include/cuvs_b200/cuvs.hpp: the reference's C++ call shape (cuvs::neighbors::ivf_pq::build/search over mdspan-like
    views, ivf_pq.hpp:1821-1828) as header-only templates over the C ABI; the example client compiles with -Wall -Werror,
    links the in-tree library and runs its GPU-free path."""
    exe = tmp_path / "cpp_client"
    _build_cpp_client(exe)
    out = subprocess.run([str(exe), "--no-gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("libcuvs_c 26."), out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_header_adaptor_search_on_the_gpu(tmp_path):
    exe = tmp_path / "cpp_client"
    _build_cpp_client(exe)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "CPP_ADAPTOR_OK" in out.stdout, out.stdout + out.stderr


def _build_bench_algo_demo(out):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = ["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", os.path.join(root, "examples", "cpp", "bench_algo_demo.cpp"),
           "-I" + os.path.join(root, "include"), "-I/usr/local/cuda/include", "-L" + os.path.join(root, "cuvs_b200", "lib"), "-lcuvs_c",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.join(root, "cuvs_b200", "lib"), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_bench_algo_wrappers_compile_and_link(tmp_path):
    """include/cuvs_b200/bench_algo.hpp: the reference harness's `algo<T>` interface (cpp/bench/ann/src/common/ann_types.hpp:83-166)
    and its cuvs_ivf_pq / cuvs_ivf_flat / brute-force wrappers (cpp/bench/ann/src/cuvs/*_wrapper.h) over the C ABI."""
    exe = tmp_path / "bench_algo_demo"
    _build_bench_algo_demo(exe)
    out = subprocess.run([str(exe), "--no-gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "algo<T> wrappers" in out.stdout, out.stdout + out.stderr
