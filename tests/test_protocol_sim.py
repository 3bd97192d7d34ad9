"""CPU test of the tensor-core kernel's synchronisation protocol: the functional simulator (tools/protocol_sim.py)
replays the producer / 4 MMA issuers / 2 epilogue sets / front-end against the real layer program with the hardware's
ONE-bit mbarrier parity semantics.  It reproduces the deadlock seen on the B200 before the 'armed stages' counter was
added (an issuer more than one ring round ahead of the producer aliases on parity) and proves the fixed protocol
free of deadlock and of premature stage consumption for every ring depth."""
import os
import sys

import pytest

from conftest import ROOT
from oracle import nerf_oracle as O

sys.path.insert(0, os.path.join(ROOT, "tools"))
from protocol_sim import simulate, simulate_pair  # noqa: E402
from test_host_logic import debug_pack  # noqa: E402


@pytest.mark.parametrize("arch", [dict(), dict(num_layers=4, hidden_size=128, num_encoding_fn_xyz=6),
                                  dict(num_layers=3, hidden_size=128, use_viewdirs=False),
                                  dict(num_layers=6, hidden_size=256, skip_step=2, num_encoding_fn_xyz=8)])
@pytest.mark.parametrize("sigma_only", [False, True])
def test_protocol_is_deadlock_free(arch, sigma_only):
    cfg = O.NetCfg(**{**O.NetCfg().__dict__, **arch})
    prog, _ = debug_pack(cfg, O.init_weights(cfg, 1), sigma_only)
    for ns in (2, 3, 4, 5, 8):
        ok, info = simulate(prog, tiles=4, NS=ns)
        assert ok, (ns, info)


def test_simulator_catches_parity_aliasing():
    cfg = O.NetCfg()
    prog, _ = debug_pack(cfg, O.init_weights(cfg, 1))
    ok, _ = simulate(prog, tiles=3, NS=5, armed_counter=False)     # the pre-fix protocol, as it failed on hardware
    assert not ok


@pytest.mark.parametrize("arch", [dict(), dict(num_layers=4, hidden_size=128, num_encoding_fn_xyz=6)])
def test_cta_pair_sharing_one_weight_stream_is_deadlock_free(arch):
    """P.cluster == 2: rank 0 multicasts every stage into both CTAs' rings, stages are released into both CTAs' w_empty
    barriers by each CTA's consuming issuer (count 2); with and without rank 1's ghost iteration, for every ring depth."""
    cfg = O.NetCfg(**{**O.NetCfg().__dict__, **arch})
    for sigma_only in (False, True):
        prog, _ = debug_pack(cfg, O.init_weights(cfg, 1), sigma_only)
        for ns in (2, 3, 5, 7):
            for ghost in (False, True):
                ok, info = simulate_pair(prog, tiles=3, NS=ns, ghost=ghost)
                assert ok, (ns, ghost, info)


def test_front_end_assisted_emission_stays_in_lockstep():
    """fe_emit (mode 1): the front-end warps answer chunk_ready[2..3] of EVERY layer with emit_done, the epilogue waits for the
    previous layer's emit_done before it arrives on chunk_ready — neither side more than one phase ahead on the one-bit parities
    (an earlier version that only synchronised on layers with an emission deadlocked / aliased on the hardware)."""
    for arch in (dict(), dict(num_layers=4, hidden_size=128, num_encoding_fn_xyz=6), dict(num_layers=3, hidden_size=128, use_viewdirs=False)):
        cfg = O.NetCfg(**{**O.NetCfg().__dict__, **arch})
        prog, _ = debug_pack(cfg, O.init_weights(cfg, 1))
        for ns in (2, 5, 7):
            for tiles in (1, 2, 4):
                ok, info = simulate(prog, tiles=tiles, NS=ns, fe_emit=True)
                assert ok, (arch, ns, tiles, info)
