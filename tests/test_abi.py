"""CPU suite: the C-ABI library loads without a GPU and exports every symbol include/srb200.h declares;
host-only entry points (shape arithmetic, argument validation, error strings) behave."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
import srb200
from srb200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "srb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    names = _declared()
    assert len(names) >= 13
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libsrb200.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "python binding table out of sync with the header"


def test_version_and_error_string():
    assert _lib.lib.srb_version() == 200
    p = _lib.ConvParams(1, 3, 8, 8, 4, 3, 3, 0, 1, 0, 0, 1, 0, 0.2, 0)  # stride 0
    ho, wo = ctypes.c_int32(), ctypes.c_int32()
    rc = _lib.lib.srb_conv_out_hw(ctypes.byref(p), ctypes.byref(ho), ctypes.byref(wo))
    assert rc == -1
    assert b"stride" in _lib.lib.srb_last_error()
    with pytest.raises(srb200.SrbError):
        _lib.check(rc)


@pytest.mark.parametrize("H,W,k,s,p,tr,op,exp", [
    (64, 64, 9, 1, 0, 0, 0, (56, 56)), (60, 60, 3, 1, 0, 0, 0, (58, 58)), (128, 128, 3, 2, 1, 0, 0, (64, 64)),
    (28, 28, 9, 4, 3, 1, 1, (112, 112)), (5, 6, 4, 2, 1, 1, 0, (10, 12)), (32, 32, 3, 1, 1, 0, 0, (32, 32)),
])
def test_out_hw(H, W, k, s, p, tr, op, exp):
    prm = _lib.ConvParams(1, 4, H, W, 4, k, k, s, p, op, tr, 1, 0, 0.2, 2)
    ho, wo = ctypes.c_int32(), ctypes.c_int32()
    assert _lib.lib.srb_conv_out_hw(ctypes.byref(prm), ctypes.byref(ho), ctypes.byref(wo)) == 0
    assert (ho.value, wo.value) == exp


def test_rejects_bad_requests():
    ho, wo = ctypes.c_int32(), ctypes.c_int32()
    bad = [
        _lib.ConvParams(1, 4, 2, 2, 4, 5, 5, 1, 0, 0, 0, 1, 0, 0.2, 0),   # kernel larger than input
        _lib.ConvParams(1, 4, 8, 8, 4, 3, 3, 1, 1, 0, 0, 1, 7, 0.2, 0),   # unknown activation
        _lib.ConvParams(1, 4, 8, 8, 4, 3, 3, 2, 1, 2, 1, 1, 0, 0.2, 0),   # out_pad >= stride
        _lib.ConvParams(1, 4, 8, 8, 4, 3, 3, 1, 1, 1, 0, 1, 0, 0.2, 0),   # out_pad on a plain conv
    ]
    for p in bad:
        assert _lib.lib.srb_conv_out_hw(ctypes.byref(p), ctypes.byref(ho), ctypes.byref(wo)) == -1
    p = _lib.ConvParams(1, 4, 8, 8, 4, 3, 3, 2, 1, 0, 1, 2, 0, 0.2, 0)        # PixelShuffle + transposed
    assert _lib.lib.srb_conv_out_hw(ctypes.byref(p), ctypes.byref(ho), ctypes.byref(wo)) == -2


def test_workspace_query_is_host_only():
    p = _lib.ConvParams(16, 64, 32, 32, 64, 3, 3, 1, 1, 0, 0, 1, 1, 0.2, 2)
    for ps in (0, 1, 2):
        assert _lib.lib.srb_conv_workspace_bytes(ctypes.byref(p), ps) >= 256


# ---- planner invariants (host only: srb_conv_describe_plan never touches the GPU) ------------------------------------
import re  # noqa: E402
lib = _lib.lib

_PLAN_SHAPES = [(n, c, h, w, co, k, p) for (n, h, w) in [(1, 7, 5), (2, 24, 24), (16, 32, 32), (128, 64, 64), (64, 128, 128), (4, 200, 300)]
                for (c, co) in [(3, 64), (64, 64), (64, 32), (32, 48), (64, 3), (256, 256), (96, 160), (64, 256), (12, 12)]
                for (k, p) in [(1, 0), (3, 1), (3, 0), (5, 0), (9, 4)]]


def test_tile_plans_respect_hardware_limits():
    """Every plan the library reports must fit one SM: <= 227 KB dynamic shared memory, <= 512 TMEM columns,
    a non-empty grid no larger than the chip can hold resident (persistent kernels), and >= 2 TMA stages when streaming."""
    buf = ctypes.create_string_buffer(1024)
    seen = 0
    for (n, c, h, w, co, k, p) in _PLAN_SHAPES:
        if h + 2 * p < k or w + 2 * p < k:
            continue
        prm = _lib.ConvParams(n, c, h, w, co, k, k, 1, p, 0, 0, 1, _lib.ACT_RELU, 0.2, _lib.MATH_AUTO)
        for pas in (0, 1, 2):
            assert lib.srb_conv_describe_plan(ctypes.byref(prm), pas, buf, 1024) == 0
            txt = buf.value.decode()
            if "no plan" in txt or "CUDA-core" in txt:
                continue
            seen += 1
            if txt.startswith("conv_rs"):  # row-stacked kernel: one CTA per SM, the whole 512-column TMEM as a ring of row blocks
                smem = int(re.search(r"smem (\d+) B", txt).group(1))
                g_, bw_, tw_ = (int(v) for v in re.search(r"(\d+) image\(s\) x (\d+) slots per M tile \((\d+) output columns", txt).groups())
                kh_, nt_, n_ = (int(v) for v in re.search(r"N = (\d+) x (\d+) = (\d+)", txt).groups())
                ring = int(re.search(r"TMEM ring (\d+) blocks", txt).group(1))
                stages = int(re.search(r"(\d+) row stages", txt).group(1))
                streams = int(re.search(r"(\d+) stream\(s\)", txt).group(1))
                stage_b = int(re.search(r"row stages x (\d+) B", txt).group(1))
                assert streams in (1, 2) and streams * stages * stage_b < smem, txt
                if streams == 2:  # each stream owns half of TMEM: a ring of >= 2 * kh blocks, else the window wraps too often
                    assert 2 * ring * int(re.search(r"N = \d+ x (\d+)", txt).group(1)) <= 512 and ring >= 2 * k, txt
                rows, grid = (int(v) for v in re.search(r"(\d+) output rows on grid (\d+)", txt).groups())
                assert smem <= 227 * 1024 and stages >= 2, txt
                assert kh_ == k and n_ == kh_ * nt_ <= 256 and nt_ % 16 == 0 and nt_ >= (co if pas == 0 else c), txt
                assert ring * nt_ <= 512 and ring > kh_, txt
                assert (g_ - 1) * bw_ + tw_ <= 128 and bw_ == tw_ + k - 1 and g_ >= 1, txt  # every valid lane inside the 128-slot M tile
                assert 1 <= grid <= 148 and grid <= rows, txt
                continue
            smem = int(re.search(r"smem (\d+) B", txt).group(1))
            tmem = int(re.search(r"tmem (\d+) cols", txt).group(1))
            gx, gy = (int(v) for v in re.search(r"grid (\d+) x (\d+)", txt).groups())
            assert smem <= 227 * 1024, txt
            assert tmem in (32, 64, 128, 256, 512), txt
            assert gx >= 1 and gy >= 1, txt
            if pas < 2:
                ctas = int(re.search(r"\((\d) CTA/SM\)", txt).group(1))
                assert gx * gy <= 148 * ctas, txt  # persistent grid: every CTA resident at once
                assert ctas == 1 or (smem <= 111 * 1024 and tmem <= 256), txt
            else:
                assert gx * gy <= 148, txt
                assert int(re.search(r"stages (\d+)", txt).group(1)) >= 1, txt
    assert seen > 200


def test_planner_queries_are_deterministic_and_cheap():
    prm = _lib.ConvParams(64, 64, 128, 128, 64, 3, 3, 1, 1, 0, 0, 1, _lib.ACT_RELU, 0.2, _lib.MATH_AUTO)
    b1, b2 = ctypes.create_string_buffer(512), ctypes.create_string_buffer(512)
    for pas in (0, 1, 2):
        lib.srb_conv_describe_plan(ctypes.byref(prm), pas, b1, 512)
        lib.srb_conv_describe_plan(ctypes.byref(prm), pas, b2, 512)
        assert b1.value == b2.value
        assert lib.srb_conv_workspace_bytes(ctypes.byref(prm), pas) == lib.srb_conv_workspace_bytes(ctypes.byref(prm), pas)


def test_band_plans_cover_every_output_pixel():
    """fprop/dgrad bands tile the output exactly: bands_h x TH and bands_w x TW cover Ho x Wo with no empty band, the halo
    box is (TH+k-1) x (TW+k-1 [+1 for the Cin<=4 tap pairs]) and an M-tile group holds every slot of the band."""
    buf = ctypes.create_string_buffer(1024)
    checked = 0
    for (n, c, h, w, co, k, p) in _PLAN_SHAPES:
        if h + 2 * p < k or w + 2 * p < k:
            continue
        ho, wo = h + 2 * p - k + 1, w + 2 * p - k + 1
        prm = _lib.ConvParams(n, c, h, w, co, k, k, 1, p, 0, 0, 1, _lib.ACT_NONE, 0.2, _lib.MATH_AUTO)
        lib.srb_conv_describe_plan(ctypes.byref(prm), 0, buf, 1024)
        txt = buf.value.decode()
        m = re.search(r"band (\d+)x(\d+) \(halo (\d+)x(\d+)\), (\d+)x(\d+) bands/img, MTB (\d+)", txt)
        if not m:
            continue
        th, tw, bh, bw, nh, nw, mtb = (int(v) for v in m.groups())
        assert nh * th >= ho and (nh - 1) * th < ho, txt
        assert nw * tw >= wo and (nw - 1) * tw < wo, txt
        assert bh == th + k - 1, txt
        assert bw in (tw + k - 1, tw + k), txt
        assert (th - 1) * bw + tw <= 128 * mtb, txt  # last real slot of the band lies inside its M tiles
        assert int(re.search(r"(\d+) bands on grid", txt).group(1)) == n * nh * nw, txt
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("seed", range(3))
def test_out_hw_matches_torch_shapes(seed):
    """srb_conv_out_hw vs the shapes ATen produces for Conv2d / ConvTranspose2d (random small cases, CPU)."""
    import random
    import torch
    import torch.nn.functional as TF
    rnd = random.Random(seed)
    for _ in range(40):
        k, s = rnd.randint(1, 5), rnd.randint(1, 3)
        p = rnd.randint(0, k - 1)
        h, w = rnd.randint(k, 12), rnd.randint(k, 12)
        tr = rnd.random() < 0.5
        op = rnd.randint(0, s - 1) if tr else 0
        x = torch.zeros(1, 2, h, w)
        if tr:
            y = TF.conv_transpose2d(x, torch.zeros(2, 3, k, k), None, s, p, op)
        else:
            y = TF.conv2d(x, torch.zeros(3, 2, k, k), None, s, p)
        prm = _lib.ConvParams(1, 2, h, w, 3, k, k, s, p, op, 1 if tr else 0, 1, _lib.ACT_NONE, 0.2, _lib.MATH_AUTO)
        ho, wo = ctypes.c_int32(), ctypes.c_int32()
        assert lib.srb_conv_out_hw(ctypes.byref(prm), ctypes.byref(ho), ctypes.byref(wo)) == 0
        assert (ho.value, wo.value) == tuple(y.shape[2:]), (k, s, p, op, h, w, tr)


def test_wgrad_plan_invariants():
    """Weight-gradient plans (tf32, Cin > 4): the band geometry covers the image, the accumulators fit TMEM for the flavour the
    planner chose (separate filter rows / rows stacked along N / rows and co blocks stacked), and the rows+co-stacked flavour keeps
    its two structural requirements: BW % 8 == 0 (no K-step straddles two rows) and N = kh * NT <= 256."""
    buf = ctypes.create_string_buffer(1024)
    flavours = {"": 0, "-rows-stacked": 0, "-rows+co-stacked": 0}
    shapes = _PLAN_SHAPES + [(64, 64, 128, 128, 64, 3, 1), (16, 64, 37, 150, 64, 3, 1), (8, 128, 40, 40, 128, 2, 0)]
    for (n, c, h, w, co, k, p) in shapes:
        if c <= 4 or c % 32 or co % 4 or h + 2 * p < k or w + 2 * p < k:
            continue
        ho, wo = h + 2 * p - k + 1, w + 2 * p - k + 1
        prm = _lib.ConvParams(n, c, h, w, co, k, k, 1, p, 0, 0, 1, _lib.ACT_NONE, 0.2, _lib.MATH_AUTO)
        assert lib.srb_conv_describe_plan(ctypes.byref(prm), 2, buf, 1024) == 0
        txt = buf.value.decode()
        m = re.match(r"tc_wgrad(\S*): band TH (\d+) TW (\d+) x(\d+) BW (\d+) BH (\d+), bands (\d+) \((\d+) per CTA\), CIB (\d+) RG (\d+) "
                     r"SG (\d+) NT (\d+), groups ci (\d+) r (\d+) co (\d+), stages (\d+), smem (\d+) B, tmem (\d+) cols, grid (\d+) x (\d+), "
                     r"ksteps (\d+)", txt)
        if not m:
            assert "no plan" in txt or "CUDA-core" in txt or "bf16" in txt, txt
            continue
        fl = m.group(1)
        th, tw, xw, bw, bh, bands, per_cta, cib, rg, sg, nt, gci, gr, gco, stages, smem, tmem, gx, gy, ksteps = (int(v) for v in m.groups()[1:])
        assert fl in flavours, txt
        flavours[fl] += 1
        assert xw * tw >= wo and (xw - 1) * tw < wo and bw >= tw + k - 1, txt
        assert sg == (k + 3) // 4 and nt % 32 == 0 and gco * nt >= co and gci * cib * 32 == c, txt
        rows = h if fl else ho  # the stacked flavours tile the INPUT rows
        assert bands == n * ((rows + th - 1) // th) * xw and gx * per_cta >= bands and gx * gy <= 148, txt
        assert ksteps == (th * bw + 7) // 8 and stages >= 2 and smem <= 227 * 1024, txt
        if fl == "-rows+co-stacked":
            assert bw % 8 == 0 and k * nt <= 256 and nt >= 64 and rg == k, txt
            assert sg * cib * k * nt <= tmem <= 512, txt
        elif fl == "-rows-stacked":
            assert k * 32 <= 256 and rg == k and bh == th, txt
            assert (nt // 32) * sg * cib * k * 32 <= tmem <= 512, txt
        else:
            assert bh == th + rg - 1 and rg * sg * cib * nt <= tmem <= 512, txt
    assert all(v > 0 for v in flavours.values()), flavours


def test_shipped_library_carries_tcgen05_tma_tmem_sass():
    """The built libsrb200.so really is a Blackwell tensor-core library: the three hot kernels contain tcgen05.mma (UTCHMMA), TMA
    tensor loads (UTMALDG), TMEM loads (LDTM), tcgen05.commit (UTCBAR) and mbarrier (SYNCS) SASS, and nothing in it is a legacy
    mma.sync (HMMA) kernel.  (`tools/sass_census.py` writes the full per-kernel table into profiles/.)"""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    so = os.path.join(os.path.dirname(_lib.__file__), "libsrb200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = {}
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for k in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "HMMA"):
                if op.startswith(k):
                    counts[cur][k] = counts[cur].get(k, 0) + 1
    assert "sm_100a" in out or "sm_100" in out
    for kern in ("k_conv_rs", "k_conv_sl", "k_tc_wgrad"):
        fn = [f for f in counts if kern + "E" in f or kern + "I" in f]
        assert fn, kern
        c = counts[fn[0]]
        for k in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS"):
            assert c.get(k, 0) > 0, (kern, k, c)
    assert sum(c.get("HMMA", 0) for c in counts.values()) == 0
