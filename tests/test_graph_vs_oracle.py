"""CPU: the product's layer graphs (graph.py) describe the same networks as the functional oracle, which is
itself pinned to the reference by tests/test_oracle_golden.py."""
import pytest
import torch

from gdn_pytorch_b200 import graph as G
from oracle import model as OM, synth
from oracle.graph_interp import run_graph
from tests.util import shapes_of

CASES = [("AutoEncoder_2", 3), ("AutoEncoder_DtoD", 1), ("AutoEncoder", 3)]


@pytest.mark.parametrize("name,cin", CASES)
@pytest.mark.parametrize("train", [False, True])
def test_graph_matches_oracle(name, cin, train):
    torch.manual_seed(0)
    H, W, B = 32, 64, 2
    sd = synth.synth_state_dict(shapes_of(name), seed=0)
    x = synth.synth_rgb(B, H, W, 0) if cin == 3 else synth.synth_depth(B, H, W, 0)
    g = G.GRAPHS[name]() if name == "AutoEncoder" else G.GRAPHS[name](cin)
    with torch.no_grad():
        T = run_graph(g, sd, x, train=train)
        ref = OM.FORWARDS[name](sd, x, istrain=True, train=train)
    for nm, r in zip(g.outputs, ref):
        got = T[nm]
        assert got.shape == r.shape, (nm, got.shape, r.shape)
        err = (got - r).abs().max().item() / (r.abs().max().item() + 1e-12)
        # train-mode nets at random init amplify fp32 summation-order noise (Tolerances)
        assert err < (5e-3 if train else 2e-4), (name, nm, err)


def test_encoder_only_outputs():
    g = G.graph_autoencoder_dtod(1)
    sd = synth.synth_state_dict(shapes_of("AutoEncoder_DtoD"), seed=0)
    x = synth.synth_depth(1, 32, 64, 0)
    with torch.no_grad():
        T = run_graph(g, sd, x, stop_after="x6")
        ref = OM.autoencoder_dtod(sd, x, encoder_only=True)
    assert "x15" not in T
    for nm, r in zip(g.encoder_outputs, ref):
        assert torch.allclose(T[nm], r, atol=1e-5)
