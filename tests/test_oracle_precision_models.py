"""The CPU emulation of the tcgen05 path's arithmetic (oracle/analysis_fp16_accumulate.py): fp16 operands with
fp32 accumulation -- today's kernel -- and with fp16 accumulators -- the next design step of DESIGN.md section 9 --
both stay inside the tolerances SURVEY 8d sets for fp16-operand kernels on the 4x128 fixture of the reference."""
import golden_util as G
from oracle import analysis_fp16_accumulate as A


def test_fp16_operand_models_within_tolerance():
    meta, a = G.load("c2_vmap_w128_s64")
    err = {}
    for mode in ("fp32", "f16op_f32acc", "f16op_f16acc"):
        p = A.render(meta, a, mode)
        err[mode] = ((p.rgbds[..., :3] - a["out_rgbds"][..., :3]).abs().mean().item(),
                     (p.rgbds[..., 3] - a["out_rgbds"][..., 3]).abs().mean().item(),
                     (p.term_probs - a["out_term_probs"]).abs().mean().item())
    assert max(err["fp32"]) < 1e-6  # the emulation harness itself reproduces the reference
    for mode in ("f16op_f32acc", "f16op_f16acc"):
        col, dep, term = err[mode]
        assert col < 2e-3 and dep < 5e-3 and term < 3e-3, (mode, err[mode])
    # fp16 accumulators cost well under 2x the error the fp16 operands already cost
    assert err["f16op_f16acc"][1] < 2 * err["f16op_f32acc"][1]
