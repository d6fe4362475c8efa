"""Host helpers of the render path (neural_graph_mapping_b200/utils.py) against the reference's contracts
(ngm/utils.py:114-138 plugin lookup, :220-251 blocked evaluation)."""
import torch

from neural_graph_mapping_b200 import utils


def test_str_to_object_scopes_and_imports():
    local_thing = object()
    assert utils.str_to_object("local_thing") is local_thing            # caller locals first
    assert utils.str_to_object("torch") is torch                        # then caller globals
    assert utils.str_to_object("torch.nn.Linear") is torch.nn.Linear    # then the dotted import path
    import neural_graph_mapping_b200.models as M

    assert utils.str_to_object("neural_graph_mapping_b200.models.NeuralField") is M.NeuralField
    assert utils.str_to_object("no.such.module.Thing") is None          # unknown names give None (callers raise)
    assert utils.str_to_object("torch.nn.NoSuchLayer") is None


def test_batched_evaluation_matches_single_call():
    x = torch.arange(23 * 3, dtype=torch.float32).view(23, 3)

    def model(b):
        keep = b[:, 0] % 2 == 0
        return b * 2, b.sum(-1), None, b[keep, 1]  # dense, dense, non-tensor, data-dependent length

    whole = model(x)
    for block in (1, 5, 23, 100):
        got = utils.batched_evaluation(model, x, block)
        assert torch.equal(got[0], whole[0]) and torch.equal(got[1], whole[1])
        assert all(v is None for v in got[2])
        assert torch.equal(got[3], whole[3])
    single = utils.batched_evaluation(lambda b: b + 1, x, 4)
    assert torch.equal(single, x + 1)
