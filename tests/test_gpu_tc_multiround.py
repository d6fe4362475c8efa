"""The persistent tcgen05 kernels with FEW CTAs (NGM_TC_MAX_CTAS): every CTA runs many rounds and
crosses field boundaries (weight-image switches), the regime of the full-size benchmark.
Runs in a subprocess because the cap is read once per process."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(240)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, torch
sys.path.insert(0, "tests")
import golden_util as G
from tests_support import make_state
import neural_graph_mapping_b200 as ngm

dev = "cuda:0"
meta, a = G.load("vmap_guided_nrgbd")
cam = ngm.Camera(**meta["camera"])
for S, Sg, F, Rr in [(64, 0, 3, 41), (16, 8, 5, 70), (128, 0, 2, 9), (24, 0, 4, 333)]:
    m = dict(meta, num_samples=S, num_samples_depth_guided=Sg)
    g = torch.Generator().manual_seed(S + F)
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1)
    near = torch.rand(F, Rr, generator=g) * 0.5 + 0.3
    far = near + 1.5
    gt = near + (far - near) * torch.rand(F, Rr, generator=g)
    jit = torch.rand(F, Rr, S, generator=g)
    jg = torch.rand(F, Rr, max(Sg, 1), generator=g)[..., :Sg]
    fid = torch.randperm(5, generator=g)[:F]
    outs = {}
    for prec in ("fp32", "fp16"):
        st = make_state(m, a, dev, prec)
        with torch.no_grad():
            outs[prec] = st._render_ijs(ijs.to(dev), a["c2ws"][0, 0].to(dev), cam, fid.to(dev), True, near.to(dev),
                                        far.to(dev), gt.to(dev) if Sg else None, jitter=jit.to(dev),
                                        jitter_guided=jg.to(dev) if Sg else None)
    torch.cuda.synchronize()
    p32, p16 = outs["fp32"], outs["fp16"]
    ec = (p16.rgbds[..., :3] - p32.rgbds[..., :3]).abs().mean().item()
    ed = (p16.rgbds[..., 3] - p32.rgbds[..., 3]).abs().mean().item()
    et = (p16.term_probs - p32.term_probs).abs().mean().item()
    print(f"S={S} Sg={Sg} F={F} R={Rr}: colour {ec:.2e} depth {ed:.2e} term {et:.2e}")
    assert ec < 2e-3 and ed < 5e-3 and et < 3e-3, (S, Sg, F, Rr, ec, ed, et)
# field forward (stage form), many tiles per CTA, several fields
from oracle import restatement as R
g = torch.Generator().manual_seed(3)
spec = R.FieldSpec("nerf", {"dim_in": 3, "num_octaves": 8}, 4, 4, 128, "no")
F, n = 3, 128 * 37 + 5
params = R.stack_params([R.init_field_params(spec, g) for _ in range(F)])
pts = torch.rand(F, n, 3, generator=g)
ref = R.fieldset_forward_vmap(pts, None, None, spec, params, R.RenderSpec(scale_mode="no"))
model = ngm.NeuralFieldSet(3, "neural_graph_mapping_b200.models.NeuralField",
                           {"encoding_type": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
                            "encoding_kwargs": {"dim_in": 3, "num_octaves": 8}, "num_layers": 4, "dim_out": 4,
                            "dim_mlp_out": 128}, 2, 10.0, 1.0, precision="fp16").to(dev)
model.all_fields_params = {k: v.to(dev) for k, v in params.items()}
model.set_vmap_fields(None)
with torch.no_grad():
    y = model(pts.to(dev), None, None, None, True)
e = (y.cpu() - ref).abs()
print("field fwd multiround: max", e.max().item(), "mean", e.mean().item(), "scale", ref.abs().max().item())
assert e.max().item() < 2e-2 * ref.abs().max().item() and e.mean().item() < 3e-3 * ref.abs().max().item()
print("MULTIROUND OK")
'''


@pytest.mark.parametrize("ctas", ["1", "2", "5"])
def test_tc_kernels_few_ctas(ctas):
    env = dict(os.environ, NGM_TC_MAX_CTAS=ctas)
    r = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=200)
    assert r.returncode == 0 and "MULTIROUND OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
