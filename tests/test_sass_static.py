"""Static checks of the BUILT library's device code (cuobjdump on libcck_b200.so; no GPU): the properties DESIGN section 3
quotes for the headline kernel must hold for the artefact that ships - register budget, no local memory, TMA / cp.async / mbarrier
instructions present, and the instruction mix of the warp-uniform hot loop that bounds its FP64 issue fraction."""
import os
import shutil
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_loop_mix as slm  # noqa: E402

SO = os.path.join(ROOT, "chance-constrained-k-cbs_b200", "libcck_b200.so")
pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")


@pytest.fixture(scope="module")
def ks(K):
    K.capi.build()
    return slm.kernels(SO)


def test_headline_kernel_resources(ks):
    for name in ("k_rollout_lin_grid<false, false>", "k_rollout_lin_grid<true, false>"):
        res = ks[name][1]
        assert res["reg"] <= 128, (name, res)          # 4 CTAs x 128 threads per SM need <= 128 registers per thread
        assert res["local"] == 0 and res["stack"] == 0, (name, res)     # no spills, no local arrays
    # the planner's expansion variants run 3 CTAs x 128 threads: <= 170 registers, no spills
    for name in ("k_rollout_lin_grid<false, true>", "k_rollout_lin_grid<true, true>"):
        res = ks[name][1]
        assert res["reg"] <= 170 and res["local"] == 0, (name, res)
    for name, (_, res) in ks.items():
        if name.startswith("k_rollout_uni2"):
            assert res["reg"] <= 255, (name, res)


def test_headline_kernel_uses_tma_cp_async_and_no_fma_contraction(ks):
    ins = slm.sass(SO, ks["k_rollout_lin_grid<false, false>"][0])
    ops = [slm.opcode(t) for _, t in ins]
    assert "UBLKCP" in ops          # cp.async.bulk (TMA 1-D bulk copy) stages the grid table
    assert "SYNCS" in ops           # mbarrier
    assert "LDGSTS" in ops          # cp.async prefetch of the next belief
    assert ops.count("ATOMS") <= 2  # ONE work-counter atomic per warp refill (round 1: one per lane)
    assert not any(o.startswith(("HMMA", "IMMA", "UTC")) for o in ops)     # 2x2 algebra: no tensor-core work to be had


def test_uniform_loop_instruction_mix(ks):
    """The warp-uniform fast section is the hot loop of the all-survive / dense-survive sweeps.  Its FP64 instructions are the
    work bit-exact results require (34 DMUL + 30 DADD + 5 DFMA, the DFMAs being the Newton steps of the one division); every
    other instruction lowers the attainable FP64 issue fraction 2f / (2f + o)."""
    ins = slm.sass(SO, ks["k_rollout_lin_grid<false, false>"][0])
    hot = slm.hot_loop(slm.loops(ins))
    assert hot is not None
    assert hot["fp64"] == 69 and hot["hist"]["DMUL"] == 34 and hot["hist"]["DADD"] == 30 and hot["hist"]["DFMA"] == 5, dict(hot["hist"])
    assert hot["n"] <= 126, hot["n"]                                   # 124 when this was written
    assert 2 * hot["fp64"] / (2 * hot["fp64"] + hot["n"] - hot["fp64"]) >= 0.70
    assert hot["hist"]["STS"] == 5 and hot["hist"].get("LDS", 0) == 0   # the stash: five STS.128, never read back in the loop
    assert hot["hist"].get("STL", 0) == 0 and hot["hist"].get("LDL", 0) == 0
