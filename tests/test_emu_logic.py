"""Kernel-source logic check without a GPU: ffmpeg-heaac_b200/csrc/heaac_kernels.cuh compiled as plain C++
(tests/emu, one 'thread' per CTA) must agree with the oracle BIT FOR BIT -- both evaluate the same IEEE
operations in the same order.  This is test infrastructure, not a product path (see tests/emu/heaac_emu.cpp);
races and barrier placement can only be checked on the GPU (tests -m gpu)."""
import numpy as np
import pytest

import pyemu
import pyoracle
from common import GOLDEN_CASES, golden_args, golden_config, load_golden
from ffmpeg_heaac_b200 import synthetic


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bias_mode", [False, True])
def test_emulated_kernels_equal_oracle_on_golden(name, bias_mode):
    d = load_golden(name)
    cfg = golden_config(d, n_streams=2, bias_mode=bias_mode)
    args = golden_args(d, n_streams=2, bias_mode=bias_mode)
    orc, emu = pyoracle.Oracle(cfg), pyemu.Emu(cfg)
    op, of = orc.synth(*args)
    ep, ef = emu.synth(*args)
    orc.close(); emu.close()
    assert np.array_equal(ef, of) and np.array_equal(ep, op)


@pytest.mark.parametrize("name", sorted(synthetic.WORKLOADS))
def test_emulated_kernels_equal_oracle_on_synthetic_workloads(name):
    cfg, a = synthetic.make_workload(name, 5, 14, seed=17, K=5)
    orc, emu = pyoracle.Oracle(cfg), pyemu.Emu(cfg)
    op, of = orc.synth(a["coeffs"], a["ics"], a["tns"], a["sbr"], a["ps"])
    ep, ef = emu.synth(a["coeffs"], a["ics"], a["tns"], a["sbr"], a["ps"])
    orc.close(); emu.close()
    assert np.isfinite(of).all()
    assert np.array_equal(ef, of) and np.array_equal(ep, op)
    # the workload must not be degenerate: signal present, no hard clipping
    assert 200 < np.sqrt((of.astype(np.float64) ** 2).mean()) < 12000
    assert (np.abs(op) >= 32767).mean() < 1e-3


@pytest.mark.parametrize("name", ["heaac_v1_stereo", "heaac_v2_ps"])
def test_frame_by_frame_equals_batched_call(name):
    """State hand-over: T frames in one call == T calls of one frame (v1: the chain keeps part of the state in "shared memory"
    for the length of a call)."""
    cfg, a = synthetic.make_workload(name, 2, 8, seed=3, K=2)
    e1, e2 = pyemu.Emu(cfg), pyemu.Emu(cfg)
    p1, f1 = e1.synth(a["coeffs"], a["ics"], a["tns"], a["sbr"], a["ps"])
    for t in range(8):
        p2, f2 = e2.synth(a["coeffs"][t:t + 1], a["ics"][t:t + 1], None, a["sbr"][t:t + 1], a["ps"][t:t + 1] if a["ps"] is not None else None)
        assert np.array_equal(f2[0], f1[t])
    e1.close(); e2.close()


def test_emulated_coupling_equals_oracle():
    """k_couple's routine (independent channel coupling + s16 epilogue) against the oracle on the coupling golden, both scalings."""
    from test_oracle_vs_ref import _cce_golden
    from common import abi
    d = _cce_golden()
    et = [int(x) for x in d["elem_types"]]
    for bias_mode in (False, True):
        cfg = abi.make_config(2, et, int(d["sr_index"]), max_frames=d["coeffs"].shape[0], out_slots=None, bias_mode=bias_mode)
        rep = lambda a: np.ascontiguousarray(np.repeat(a[:, None], 2, axis=1))
        co = rep(d["coeffs"]) * (np.float32(1.0 / 32768.0) if bias_mode else np.float32(1.0))
        orc, emu = pyoracle.Oracle(cfg), pyemu.Emu(cfg)
        op, of = orc.synth(co, rep(d["ics"]), rep(d["tns"]), None, None, cce=rep(d["cce"]))
        ep, ef = emu.synth(co, rep(d["ics"]), rep(d["tns"]), None, None, cce=rep(d["cce"]))
        orc.close(); emu.close()
        assert np.array_equal(ef, of) and np.array_equal(ep, op)
