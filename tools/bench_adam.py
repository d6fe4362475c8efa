"""Secondary measurement: the optimizer half of one mapping iteration at BASELINE config 4's per-GPU shard (32 active
of 256 fields, 4-layer x 128 MLP + NeRF-8): the reference's sequence on the GPU -- gather the moments
(_set_vmap_fields, ngm/run_mapping.py:679-707), torch.optim.Adam.step(), scatter parameters and moments back
(_update_step, :1191-1221) -- against one ngm_adam_step launch on the full tables."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from neural_graph_mapping_b200 import optim  # noqa: E402

dev = "cuda:0"
F_ALL, F_ACT, L, W, E = 256, 32, 4, 128, 48
g = torch.Generator().manual_seed(0)
shapes = {}
for i, (di, do) in enumerate(zip([E] + [W] * L, [W] * L + [4])):
    shapes[f"_linears.{i}.weight"], shapes[f"_linears.{i}.bias"] = (do, di), (do,)
params = {k: torch.randn(F_ALL, *s, generator=g).to(dev) for k, s in shapes.items()}
ids = torch.randperm(F_ALL, generator=g)[:F_ACT].to(dev)
grads = {k: torch.randn(F_ACT, *s, generator=g).to(dev) * 1e-3 for k, s in shapes.items()}
lr, eps, wd = 1e-3, 1e-15, 1e-5
elems = sum(v[0].numel() for v in params.values()) * F_ACT


def timed(fn, reps=30):
    ts = []
    for i in range(reps + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


# the reference's sequence
ref_all = {k: v.clone() for k, v in params.items()}
ref_state = {k: {"step": torch.tensor(1.0), "exp_avg": torch.zeros_like(v), "exp_avg_sq": torch.zeros_like(v)}
             for k, v in ref_all.items()}
dummy = torch.zeros((), device=dev, requires_grad=True)
opt = torch.optim.Adam([dummy], lr=lr, eps=eps, weight_decay=wd)


def reference_sequence():
    with torch.no_grad():
        vm = {k: v[ids] for k, v in ref_all.items()}  # models.py:274-276
    for p in vm.values():
        p.requires_grad_()
    for k, p in vm.items():  # run_mapping.py:685-705
        opt.state[p] = {"step": ref_state[k]["step"], "exp_avg": ref_state[k]["exp_avg"][ids],
                        "exp_avg_sq": ref_state[k]["exp_avg_sq"][ids]}
        p.grad = grads[k]
    for old in list(opt.state):
        if all(old is not p for p in vm.values()):
            del opt.state[old]
    opt.param_groups[0]["params"] = list(vm.values())
    opt.step()
    with torch.no_grad():  # :1195-1215
        for k, p in vm.items():
            ref_all[k][ids] = p
            ref_state[k]["exp_avg"][ids] = opt.state[p]["exp_avg"]
            ref_state[k]["exp_avg_sq"][ids] = opt.state[p]["exp_avg_sq"]


ours_all = {k: v.clone() for k, v in params.items()}
ours_state = optim.new_optim_state(ours_all)
vm_ours = {k: v[ids].clone().requires_grad_(True) for k, v in ours_all.items()}
for k, p in vm_ours.items():
    p.grad = grads[k]


def ours():
    optim.adam_step(ours_all, vm_ours, ours_state, ids, lr, eps, wd)


t_ref, t_ours = timed(reference_sequence), timed(ours)
bytes_moved = elems * (16 + 16 + 4)
print(json.dumps({"workload": f"Adam on {F_ACT} of {F_ALL} fields, {L}x{W} MLP + NeRF-8 ({elems} elements)",
                  "reference_sequence_ms": round(t_ref, 4), "ngm_adam_step_ms": round(t_ours, 4),
                  "speedup": round(t_ref / t_ours, 1), "ngm_adam_step_GBps": round(bytes_moved / t_ours / 1e6, 1)}))
