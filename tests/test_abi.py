"""CPU tests of the boundary: the library loads and exports every symbol include/kcf_b200.h declares.
No compute call is made (there is no GPU here); kcf_init must fail loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

from kcftools_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "kcf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kcf_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    names = _declared()
    assert len(names) >= 20
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kcf_b200.h but not exported by libkcfgpu.so"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype in kcftools_b200/_lib.py"
    assert set(_lib.SYMBOLS) == set(names)


def test_struct_sizes_match_header():
    assert _lib.RESULT_DTYPE.itemsize == 48
    assert _lib.RESULT_DTYPE.fields["kmer_count_sum"][1] == 32 and _lib.RESULT_DTYPE.fields["score"][1] == 40
    assert _lib.WINDOW_DTYPE.itemsize == 8 and _lib.SEGMENT_DTYPE.itemsize == 12
    assert C.sizeof(_lib.DbInfo) == 128 and _lib.DbInfo.load_phase_s.offset == 96  # kcf_db_info_t: load_phase_s[4] appended in round 2
    assert C.sizeof(_lib.HostSeq) == 32  # kcf_host_seq_t
    assert _lib.CELL_DTYPE.itemsize == 40 and _lib.CELL_DTYPE.fields["kmer_count"][1] == 24 and _lib.CELL_DTYPE.fields["score"][1] == 32  # kcf_cell_t


def test_version_string():
    assert b"sm_100a" in _lib.load().kcf_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from kcftools_b200.api import Context, KcfError
    with pytest.raises(KcfError) as e:
        Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "kcftools_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.lower() or f == "README.md", f"{f} mentions the oracle"


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """the boundary is a C ABI: the header must compile as C99 (no C++ types) and a plain C program must link the library;
    without a device kcf_init fails loudly (no fallback), with one it succeeds"""
    import subprocess
    hdr = os.path.join(ROOT, "include", "kcf_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "kcf_b200.h"\n'
                   'int main(void) { kcf_ctx *c = NULL; int rc = kcf_init(0, &c);\n'
                   '  printf("%s|%d|%zu|%zu|%zu|%zu|%zu|%zu\\n", kcf_version(), rc, sizeof(kcf_result_t), sizeof(kcf_cell_t), sizeof(kcf_window_t), sizeof(kcf_segment_t), sizeof(kcf_db_info_t), sizeof(kcf_host_seq_t));\n'
                   '  if (rc != KCF_OK) { printf("%s\\n", kcf_last_error(NULL)); return 0; }\n  kcf_shutdown(c); return 0; }\n')
    exe = tmp_path / "abi"
    lib = os.path.join(ROOT, "kcftools_b200")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src), "-L", lib, "-lkcfgpu", f"-Wl,-rpath,{lib}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    ver, rc, *sizes = out[0].split("|")
    assert "sm_100a" in ver and [int(x) for x in sizes] == [48, 40, 8, 12, 128, 32]
    import torch
    if torch.cuda.is_available():
        assert int(rc) == 0
    else:
        assert int(rc) < 0 and "no CPU fallback" in out[1]
