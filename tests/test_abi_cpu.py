"""CPU: the C ABI boundary.  libyvb200.so loads without a GPU and exports every function include/yvb200.h declares,
the ctypes binding lists exactly those functions, and the ctypes mirrors of the ABI structs have the size and field
offsets the C compiler gives the header's structs (checked by compiling a probe with gcc).  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "yvb200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return set(re.findall(r"\b(yv_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    from yvb200 import lib
    if not lib.available():
        pytest.fail("libyvb200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    declared = _declared()
    assert len(declared) >= 25
    assert set(lib.SYMBOLS) == declared, (sorted(declared - set(lib.SYMBOLS)), sorted(set(lib.SYMBOLS) - declared))
    cdll = C.CDLL(lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(cdll, name), name
    h = lib.load()
    assert h.yv_version() >= 100
    assert isinstance(h.yv_last_error(), bytes)
    assert h.yv_launch_count() >= 0
    # argument errors come back as a non-zero status with a message, without touching a device
    assert h.yv_gemm(None, None) != 0 and b"NULL" in h.yv_last_error()
    assert h.yv_gemm_set_variant(C.c_int(7)) != 0 and b"variant" in h.yv_last_error()
    assert h.yv_gemm_set_variant(C.c_int(0)) == 0
    # fused attention: supported head sizes, workspace query and argument checks (no device work)
    assert h.yv_attn_supported(C.c_int32(128), C.c_int32(3)) == 1 and h.yv_attn_supported(C.c_int32(96), C.c_int32(3)) == 0
    assert lib.attn_bwd_workspace_bytes(8, 8, 128, 288, 288) == 3 * 8 * 288 * 2 * 1024 * 4
    assert h.yv_attn_fwd(None, None) != 0 and b"NULL" in h.yv_last_error()
    bad = lib.YvAttnFwd()
    bad.pairs, bad.heads, bad.dh, bad.passes = 1, 1, 96, 3
    assert h.yv_attn_fwd(C.byref(bad), None) != 0 and b"head size" in h.yv_last_error()


def test_ctypes_structs_match_the_header_layout(tmp_path):
    import struct
    from yvb200 import lib
    probe = tmp_path / "probe.c"
    fields = {
        "YvOperand": ["ptr", "inner", "rows", "ld", "nb0", "sb0", "nb1", "sb1", "plane_stride", "mn_major"],
        "YvGemm": ["M", "N", "K", "passes", "a", "b", "alpha", "act", "bias", "aux_out", "aux_in", "residual", "out32",
                   "ld_out", "out_sb0", "out_sb1", "out_planes", "ld_pl", "pl_sb0", "pl_sb1", "pl_plane_stride", "drop_p",
                   "drop_site", "rng", "out32_zeroed"],
        "YvSplitSeg": ["src", "dst_off", "numel", "first_blk"],
        "YvHeadView": ["ptr", "ld", "plane_stride", "pair_stride", "rows"],
        "YvAttnFwd": ["pairs", "heads", "dh", "passes", "q", "k", "v", "mask", "scale", "drop_p", "drop_site", "rng",
                      "out_planes", "ld_out", "out_plane_stride", "out32", "ld_out32", "lse"],
        "YvAttnBwd": ["pairs", "heads", "dh", "passes", "q", "k", "v", "dout", "out", "mask", "scale", "drop_p", "drop_site",
                      "rng", "lse", "dq", "dk", "dv", "workspace", "workspace_bytes", "tickets"],
        "YvAdamSeg": ["p", "g", "m", "v", "plane_hi", "plane_lo", "numel", "first_blk", "weight_decay"],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for st, fs in fields.items():
        lines.append(f'  printf("{st} %zu", sizeof({st}));')
        for f in fs:
            lines.append(f'  printf(" %zu", offsetof({st}, {f}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    probe.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(probe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    mirrors = {"YvOperand": lib.YvOperand, "YvGemm": lib.YvGemm, "YvSplitSeg": lib.YvSplitSeg,
               "YvHeadView": lib.YvHeadView, "YvAttnFwd": lib.YvAttnFwd, "YvAttnBwd": lib.YvAttnBwd}
    for line in out:
        name, size, *offs = line.split()
        if name == "YvAdamSeg":        # yvb200/optim.py packs these rows with struct.pack("<QQQQQQqqfi", ...)
            assert struct.calcsize("<QQQQQQqqfi") == int(size)
            assert [int(o) for o in offs] == [0, 8, 16, 24, 32, 40, 48, 56, 64]
            continue
        m = mirrors[name]
        assert C.sizeof(m) == int(size), (name, C.sizeof(m), size)
        for f, off in zip(fields[name], offs):
            assert getattr(m, f).offset == int(off), (name, f, getattr(m, f).offset, off)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """The CUDA path has no fallback: without libyvb200.so, loading raises and names the build command."""
    from yvb200 import lib, ops
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "libyvb200.so"))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        lib.load()
    # and the operator layer refuses host tensors instead of computing something else
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.rt("cpu")
    import torch
    from yvb200 import masking
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        masking.randomize_regions(torch.zeros(1, 2, 4), torch.zeros(1, 2, 3), torch.ones(1, 2, dtype=torch.long))
