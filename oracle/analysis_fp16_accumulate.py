"""Precision study (CPU emulation, no GPU needed): what would fp16 ACCUMULATORS cost the tcgen05 render path?

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The fused kernel keeps fp32 accumulators in TMEM; fp16
accumulators (read back with ``tcgen05.ld ... .pack::16b``, two columns per register) would halve the epilogue's
registers per load and drop its fp32->fp16 convert.  This script renders golden fixtures through the oracle with
the per-field MLP evaluated under three arithmetic models and reports the error against the reference's outputs:

  fp32          the reference arithmetic
  f16op_f32acc  fp16 operands, fp32 accumulation, activations rounded to fp16 between layers, bias + ReLU on the
                packed fp16 value (what ``tc_kernel`` does today)
  f16op_f16acc  same, but the accumulator is rounded to fp16 after every K=16 MMA step

    python oracle/analysis_fp16_accumulate.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as G  # noqa: E402
from oracle import restatement as R  # noqa: E402


def make_forward(mode):
    def field_forward(points, spec, params):
        h = R.encode(points, spec, params)
        if mode == "fp32":
            for i in range(spec.num_layers + 1):
                h = torch.nn.functional.linear(h, params[f"_linears.{i}.weight"], params[f"_linears.{i}.bias"])
                if i < spec.num_layers:
                    h = torch.relu(h)
            return h
        h = h.half()
        for i in range(spec.num_layers + 1):
            w = params[f"_linears.{i}.weight"].half()
            b = params[f"_linears.{i}.bias"]
            k_dim = h.shape[-1]
            if mode == "f16op_f32acc":
                acc = h.float() @ w.float().T
            else:
                acc = torch.zeros(*h.shape[:-1], w.shape[0], dtype=torch.float16)
                for k0 in range(0, k_dim, 16):
                    part = h[..., k0:k0 + 16].float() @ w[:, k0:k0 + 16].float().T
                    acc = (acc.float() + part).half()
            if i == spec.num_layers:
                return acc.float() + b  # last layer: fp32 bias add in the compositor's front end
            h = torch.relu(acc.half() + b.half())  # packed half2 bias + ReLU epilogue
        raise AssertionError

    return field_forward


def render(meta, a, mode):
    fs, rs, cam = G.field_spec(meta["field_kwargs"]), G.render_spec(meta), G.camera_spec(meta["camera"])
    g = lambda k: a[k] if k in a else None  # noqa: E731
    orig = R.field_forward
    R.field_forward = make_forward(mode)
    try:
        with torch.no_grad():
            return R.render_rays(a["ijs"], a["c2ws"], cam, rs, fs, G.params(a), a["positions"], a["orientations"],
                                 field_ids=a["field_ids"], use_vmap=True, near_distances=g("near"), far_distances=g("far"),
                                 gt_distances=g("gt"), jitter=g("jitter"), jitter_guided=g("jitter_guided"))
    finally:
        R.field_forward = orig


def main():
    torch.set_num_threads(8)
    print(f"{'fixture':22s} {'arithmetic':14s} {'colour L1':>10s} {'depth L1':>10s} {'term L1':>10s} {'PSNR dB':>8s}")
    for name in ("c2_vmap_w128_s64", "c1_vmap_256x32", "trained_sphere"):
        meta, a = G.load(name)
        a = dict(a)
        if name == "trained_sphere":
            H, W = meta["image_hw"]
            a["jitter"] = torch.rand(1, H * W, meta["num_samples"], generator=torch.Generator().manual_seed(meta["jitter_seed"]))
        for mode in ("fp32", "f16op_f32acc", "f16op_f16acc"):
            p = render(meta, a, mode)
            col = (p.rgbds[..., :3] - a["out_rgbds"][..., :3]).abs().mean().item()
            dep = (p.rgbds[..., 3] - a["out_rgbds"][..., 3]).abs().mean().item()
            term = (p.term_probs - a["out_term_probs"]).abs().mean().item() if "out_term_probs" in a else float("nan")
            psnr = ""
            if name == "trained_sphere":
                psnr = "%.3f" % R.psnr(p.rgbds[0, :, :3].reshape(H, W, 3), a["gt_rgb"][0].reshape(H, W, 3))
            print(f"{name:22s} {mode:14s} {col:10.2e} {dep:10.2e} {term:10.2e} {psnr:>8s}")


if __name__ == "__main__":
    main()
