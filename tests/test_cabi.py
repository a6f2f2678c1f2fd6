"""CPU: the C-ABI library loads without a GPU, exports every symbol include/msig.h declares, the
ctypes table covers them, and it fails loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import pytest

import msig_b200  # noqa: F401
from msig_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "msig.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(msig_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    dll = ctypes.CDLL(lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, f"declared in msig.h but not exported: {missing}"


def test_ctypes_table_matches_header():
    declared = set(_declared())
    table = set(lib.exported_symbols())
    assert table <= declared, f"bound but not declared: {sorted(table - declared)}"
    assert declared <= table, f"declared but not bound: {sorted(declared - table)}"


def test_version_and_loud_failure_without_gpu():
    import torch
    l = lib.load()
    assert l.msig_version() >= 100
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU path"):
        lib.init(0)


def test_modules_refuse_cpu_tensors():
    import torch
    from msig_b200 import model
    G = model.StyleCycleGANGenerator()
    with pytest.raises(RuntimeError, match="no CPU path"):
        G(torch.zeros(1, 3, 64, 64), torch.zeros(1, 256))
    D = model.MultiDomainDiscriminator(num_domains=3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        D(torch.zeros(1, 3, 64, 64), None)


def test_host_only_entry_points_run_without_gpu():
    """Pure host-side entry points (sizes, tables, row counts) need no device: they are what a binding in
    another language would call first, and they must agree with the layouts documented in msig.h."""
    l = lib.load()
    # packed-weight sizes: OIHW 256x256x3x3 -> [256][9][256] bf16
    d = lib.WpackDesc(lib.WPACK_FWD, 256, 256, 3, 3)
    assert l.msig_wpack_elems(ctypes.byref(d)) == 256 * 9 * 256
    # row-patch layout: 64 x 3 x 7 x 7 -> [64][7][64]; row-fold layout: 3 x 64 x 7 x 7 -> [7][32][64]
    assert l.msig_wpack_elems(ctypes.byref(lib.WpackDesc(lib.WPACK_ROWPATCH, 64, 3, 7, 7))) == 64 * 7 * 64
    assert l.msig_wpack_elems(ctypes.byref(lib.WpackDesc(lib.WPACK_ROWFOLD, 3, 64, 7, 7))) == 7 * 32 * 64
    assert l.msig_wpack_elems(ctypes.byref(lib.WpackDesc(lib.WPACK_ROWFOLD, 8, 64, 7, 7))) == 0   # > 4 channels: refused
    # epilogue-statistics rows per image: 4 per 128-pixel tile (64x64 plane -> 32 tiles; 4 phases of 128x128)
    assert l.msig_epilogue_stats_rows(64, 64, 1) == 32 * 4
    assert l.msig_epilogue_stats_rows(128, 128, 4) == 128 * 4 * 4
    # statistics workspace covers partials + coefficients + tickets for any chunking the kernels choose
    n, hw, c = 32, 4096, 256
    ws = l.msig_in_stats_workspace(n, hw, c)
    assert ws >= (n * 2 * c + n) * 4 and ws % 4 == 0
    # pack table: two jobs -> prefix sums over their element counts (host memory only)
    jobs = (lib.WpackJob * 2)()
    jobs[0] = lib.WpackJob(lib.WpackDesc(lib.WPACK_FWD, 64, 64, 3, 3), 0, 0, 0x1000, 0x2000, 0)
    jobs[1] = lib.WpackJob(lib.WpackDesc(-1, 0, 0, 0, 0), 0, 0, 0x3000, 0x4000, 512)
    buf = ctypes.create_string_buffer(l.msig_wpack_table_bytes(2))
    total, tiles = ctypes.c_int64(0), ctypes.c_int64(0)
    assert l.msig_wpack_table_build(jobs, 2, buf, ctypes.byref(total), ctypes.byref(tiles)) == 0
    # the 64x64x3x3 FWD pack (576 elements per output channel) goes through the tiled kernel, one tile per
    # output channel; the fp32 copy stays with the generic kernel
    assert total.value == 512 and tiles.value == 64
    jobs[0] = lib.WpackJob(lib.WpackDesc(lib.WPACK_CONVT_FWD, 128, 256, 4, 4), 0, 0, 0x1000, 0x2000, 0)
    assert l.msig_wpack_table_build(jobs, 2, buf, ctypes.byref(total), ctypes.byref(tiles)) == 0
    assert total.value == 128 * 256 * 16 + 512 and tiles.value == 0
    jobs[1].src = 0                                       # null source: refused with a message
    assert l.msig_wpack_table_build(jobs, 2, buf, ctypes.byref(total), ctypes.byref(tiles)) != 0
    assert b"null pointer" in l.msig_last_error()


def test_fusion_threshold_and_graph_defaults():
    from msig_b200 import ops
    # residual-block convs (9 taps x 256 ch = 36 K blocks) fuse their statistics; the 128->64 transposed conv
    # (4 taps x 128 ch = 8 K blocks) does not
    assert ops.epi_fusable(9, 256) and not ops.epi_fusable(4, 128)
    es = ops.EpiStats.__new__(ops.EpiStats)               # layout arithmetic only (no allocation)
    assert ops.epi_stats_rows(2, 100, 64, "cpu") is None  # 100 pixels: a 128-row tile would span two images
