"""CPU: the drop-in boundary.  libbfc_b200.so loads without a GPU and exports every function and global that
include/*.h declares (the reference's bfc.h / bbf.h / htab.h / bseq.h surface plus the batch ABI bfc_b200.h); and
with no CUDA device the product fails loudly instead of computing anything on the CPU."""
import ctypes as C
import glob
import os
import re

import pytest

import bfc_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PREFIXES = ("bfc_", "bfcg_", "bseq_", "kt_", "cputime", "realtime")


def declared_functions():
    names = {}
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        if os.path.basename(h) == "kmer.h":
            continue  # inline arithmetic (BFC_HD), nothing to export
        text = re.sub(r"/\*.*?\*/", " ", open(h).read(), flags=re.S)
        text = re.sub(r"//[^\n]*", " ", text)
        text = re.sub(r"^\s*#.*$", " ", text, flags=re.M)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
            n = m.group(1)
            if n.startswith(PREFIXES) and not n.endswith("_t"):
                names.setdefault(n, os.path.basename(h))
    return names


def test_every_declared_symbol_is_exported():
    L = C.CDLL(bfc_b200.lib_path())
    decl = declared_functions()
    assert len(decl) > 40, decl  # the parser found the surface
    for must in ("bfc_bf_init", "bfc_bf_insert", "bfc_bf_get", "bfc_bf_destroy", "bfc_ch_init", "bfc_ch_insert", "bfc_ch_get",
                 "bfc_ch_kmer_occ", "bfc_ch_count", "bfc_ch_hist", "bfc_ch_dump", "bfc_ch_restore", "bfc_ch_get_k", "bfc_ch_destroy",
                 "bfc_count", "bfc_correct", "kt_for", "kt_pipeline", "cputime", "realtime", "bseq_open", "bseq_close", "bseq_read",
                 "bfcg_count_batch", "bfcg_correct_batch", "bfcg_trim_batch", "bfcg_enum_records", "bfcg_count_records",
                 "bfcg_count_record_runs"):
        assert must in decl, must
    missing = [f"{n} ({h})" for n, h in sorted(decl.items()) if not hasattr(L, n)]
    assert not missing, missing
    for var, ctype in (("bfc_verbose", C.c_int), ("bfc_real_time", C.c_double), ("seq_nt6_table", C.c_ubyte * 256)):
        ctype.in_dll(L, var)
    C.c_uint64.in_dll(L, "bfc_kmer_null")
    assert list((C.c_ubyte * 256).in_dll(L, "seq_nt6_table"))[ord("A")] == 1  # reference bseq.c:9-26


def test_no_gpu_means_loud_failure_not_a_cpu_path(capfd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = bfc_b200.lib()
    assert L.bfcg_device_count() == 0
    assert not L.bfc_bf_init(24, 4)                      # NULL, as on an allocation failure in the reference
    assert not L.bfc_ch_init(31, 20)
    opt = bfc_b200.make_opt(k=31, bf_shift=24)
    b = bfc_b200.api.Batch()
    bf = bfc_b200.api.BF(24, 4, None)
    assert L.bfcg_count_batch(C.byref(opt), C.byref(bf), None, None, C.byref(b), None) != 0
    assert L.bfcg_sync() != 0
    err = capfd.readouterr().err
    assert "[E::" in err and "no CUDA device" in err
    assert b"no CUDA device" in L.bfcg_last_error()
