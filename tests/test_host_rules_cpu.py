"""Host logic of libgsd_b200.so that needs no GPU: which kernel / configuration every conv3x3 layer of the U-Net gets
(csrc/conv_host.h: halo-resident vs tap-streaming kernel, UMMA width, resident weights, CTA pairs, ring depths, grid)
and the chunk schedule of the host pipeline (csrc/plan.cu).  The expected batch-64 configurations are the kernels of
the committed ncu launch list (profiles/r2_fwd_launches.md, profiles/r2_ncu_full_convs.md), i.e. what was measured on the B200."""
import ctypes as C
import itertools

import pytest

from gelslim_depth_b200._lib import lib

SMS = 148
SMEM_MAX = 227 * 1024
KEYS = ("halo", "bn", "mt", "wres", "nepi", "cta2", "na", "nb", "smem", "grid", "th", "tw")


def plan(B, H, W, C0, C1, Cout, sms=SMS, env=None):
    out = (C.c_int * 12)()
    rc = lib.gsd_debug_plan_conv3x3(B, H, W, C0, C1, Cout, sms, out)
    assert rc == 0, lib.gsd_last_error().decode()
    return dict(zip(KEYS, out))


def unet_layers(H, W, dims=(64, 128, 256, 512, 1024)):
    """(name, H, W, C0, C1, Cout) of the 3x3 convs after the first one (unet.py:60-88)."""
    hs, ws = [H], [W]
    for _ in dims[1:]:
        hs.append(hs[-1] // 2)
        ws.append(ws[-1] // 2)
    layers = [("inc.3", hs[0], ws[0], dims[0], 0, dims[0])]
    for i in range(1, len(dims)):
        layers.append((f"down.{i - 1}.0", hs[i], ws[i], dims[i - 1], 0, dims[i]))
        layers.append((f"down.{i - 1}.3", hs[i], ws[i], dims[i], 0, dims[i]))
    for j, i in enumerate(range(len(dims) - 2, -1, -1)):
        layers.append((f"up.{j}.conv.0", hs[i], ws[i], dims[i], dims[i], dims[i]))
        layers.append((f"up.{j}.conv.3", hs[i], ws[i], dims[i], 0, dims[i]))
    return layers


# (halo kernel, N, M tiles, resident weights, epilogue warps, CTA pair) measured at batch 64 -- profiles/r2_ncu_full_convs.md
EXPECTED_B64 = {
    "inc.3": (1, 64, 1, 1, 8, 1), "down.0.0": (1, 128, 2, 0, 8, 1), "down.0.3": (1, 128, 2, 0, 8, 1),
    "down.1.0": (1, 128, 2, 0, 8, 1), "down.1.3": (1, 128, 2, 0, 8, 1), "down.2.0": (1, 256, 1, 0, 8, 1),
    "down.2.3": (1, 256, 1, 0, 8, 1), "down.3.0": (0, 256, 1, 0, 8, 1), "down.3.3": (0, 256, 1, 0, 8, 1),
    "up.0.conv.0": (1, 256, 1, 0, 8, 1), "up.0.conv.3": (1, 256, 1, 0, 8, 1), "up.1.conv.0": (1, 256, 1, 0, 8, 1),
    "up.1.conv.3": (1, 128, 2, 0, 8, 1), "up.2.conv.0": (1, 128, 2, 0, 8, 1), "up.2.conv.3": (1, 128, 2, 0, 8, 1),
    "up.3.conv.0": (1, 64, 2, 0, 8, 1), "up.3.conv.3": (1, 64, 1, 1, 8, 1),
}


def test_batch64_layer_configurations_match_the_measured_launch_list(monkeypatch):
    for var in ("GSD_CTA2", "GSD_NO_BN256", "GSD_BN256_ALL", "GSD_WRES0", "GSD_NA", "GSD_NB_MAX", "GSD_NO_HALO", "GSD_FORCE_TILE"):
        monkeypatch.delenv(var, raising=False)
    got = {}
    for name, h, w, c0, c1, co in unet_layers(320, 427):
        p = plan(64, h, w, c0, c1, co)
        got[name] = (p["halo"], p["bn"], p["mt"], p["wres"], p["nepi"], p["cta2"])
    assert got == EXPECTED_B64


def test_cta_pair_rule(monkeypatch):
    """Streamed-weight halo layers run as CTA pairs at every batch size (measured: batch 1-3 gain 3-6 %, profiles/r2_latency_b1.md),
    resident-weight 64-channel layers from ~100 tiles per SM on (batch >= 16 at 320x427: 5-9 % with the 4-deep accumulator ring);
    the tap-streaming kernel pairs only with a full wave of pair items."""
    monkeypatch.delenv("GSD_CTA2", raising=False)
    for B in (1, 2, 4, 64):
        for name, h, w, c0, c1, co in unet_layers(320, 427):
            p = plan(B, h, w, c0, c1, co)
            if p["halo"]:
                m_tiles = ((w + 7) // 8) * ((h + 15) // 16) * B
                assert p["cta2"] == ((1 if m_tiles >= 100 * SMS else 0) if p["wres"] else 1), (B, name, p)
            elif p["cta2"]:
                m_tiles = ((w + p["tw"] - 1) // p["tw"]) * ((h + p["th"] - 1) // p["th"]) * B
                assert (m_tiles + 1) // 2 * (co // p["bn"]) >= SMS // 2, (B, name, p)
    # the switch: 0 = never, 2 = every 64-channel-block conv (what the parity tests use to reach the pair kernels)
    monkeypatch.setenv("GSD_CTA2", "0")
    assert all(plan(64, h, w, c0, c1, co)["cta2"] == 0 for _, h, w, c0, c1, co in unet_layers(320, 427))
    monkeypatch.setenv("GSD_CTA2", "2")
    assert all(plan(1, h, w, c0, c1, co)["cta2"] == 1 for _, h, w, c0, c1, co in unet_layers(320, 427))


@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_every_planned_launch_fits_the_sm(mode, monkeypatch):
    """Invariants over batches 1..256 and both network geometries (G2 320x427, G3 160x213), for every pairing mode."""
    monkeypatch.setenv("GSD_CTA2", mode)
    for (H, W), B in itertools.product(((320, 427), (160, 213), (48, 59)), (1, 2, 3, 5, 8, 16, 64, 256)):
        for name, h, w, c0, c1, co in unet_layers(H, W):
            p = plan(B, h, w, c0, c1, co)
            tag = (mode, H, W, B, name, p)
            assert 1 <= p["grid"] <= SMS, tag
            assert p["bn"] in (64, 128, 256) and co % p["bn"] == 0, tag
            assert p["th"] * p["tw"] == 128, tag
            if p["cta2"]:
                assert p["grid"] % 2 == 0 and p["grid"] >= 2, tag             # whole clusters of two CTAs
                assert mode == "2" or not p["wres"] or B >= 16, tag                   # resident-weight layers pair only at large batches
            if p["halo"]:
                assert 0 < p["smem"] <= SMEM_MAX, tag
                assert p["na"] >= 2, tag
                assert p["wres"] or p["nb"] >= 3, tag
                assert 2 * p["mt"] * p["bn"] <= 512, tag                        # double-buffered accumulators fit TMEM
                assert p["nepi"] in (4, 8), tag


def test_odd_sm_count_never_pairs(monkeypatch):
    monkeypatch.setenv("GSD_CTA2", "2")
    assert plan(64, 160, 213, 128, 0, 128, sms=147)["cta2"] == 0


def schedule(batch, chunk, first=0, last=0):
    out = (C.c_int * 1024)()
    n = lib.gsd_debug_chunk_schedule(batch, chunk, first, last, out, 1024)
    assert n >= 1, lib.gsd_last_error().decode()
    return list(out)[:n]


def test_chunk_schedule_covers_every_frame_once():
    assert schedule(64, 16, 8, 8) == [8, 16, 16, 16, 8]           # bench.py's blocking-call schedule
    assert schedule(64, 64) == [64]
    assert schedule(5, 2) == [2, 2, 1]
    assert schedule(5, 2, 1, 1) == [1, 2, 1, 1]                    # tests/test_gpu_parity.py's ramp case
    for batch, chunk, first, last in itertools.product((1, 2, 5, 16, 63, 64, 257), (1, 2, 4, 16, 64, 300), (0, 1, 2, 8), (0, 1, 2, 8)):
        s = schedule(batch, chunk, first, last)
        assert sum(s) == batch and all(1 <= c <= min(chunk, batch) for c in s), (batch, chunk, first, last, s)
