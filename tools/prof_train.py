"""One launch of each training-side kernel at the reference's default training shape, for ncu:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'adam|target|observed|composite_bwd|encode' --csv --log-file gpurun_out/train_kernels.csv \
        python tools/prof_train.py
Adam: 32 active of 256 fields, 4x128 MLP + NeRF-8; target sampling: 32 fields x 512 rays over 100 keyframes of
640x480; observed fields: 500 pixels against 256 fields.  (Numbers printed under ncu are never bench values.)"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import neural_graph_mapping_b200 as ngm  # noqa: E402
from neural_graph_mapping_b200 import optim, targets  # noqa: E402

dev = "cuda:0"
g = torch.Generator().manual_seed(0)
F_ALL, F_ACT, L, W, E = 256, 32, 4, 128, 48
shapes = {}
for i, (di, do) in enumerate(zip([E] + [W] * L, [W] * L + [4])):
    shapes[f"_linears.{i}.weight"], shapes[f"_linears.{i}.bias"] = (do, di), (do,)
params = {k: torch.randn(F_ALL, *s, generator=g).to(dev) for k, s in shapes.items()}
ids = torch.randperm(F_ALL, generator=g)[:F_ACT].to(dev)
state = optim.new_optim_state(params)
vm = {k: v[ids].clone().requires_grad_(True) for k, v in params.items()}
for k, p in vm.items():
    p.grad = torch.randn(p.shape, generator=g).to(dev) * 1e-3

K = 100
cam = ngm.Camera(640, 480, 554.2562584220408, 554.2562584220408, 319.5, 239.5)
drv = types.SimpleNamespace()
drv._device, drv._camera, drv._field_radius = dev, cam, 1.0
drv._num_train_fields, drv._num_rays_per_field = 32, 512
pos = torch.randn(F_ALL, 3, generator=g) * torch.tensor([1.0, 0.7, 0.5]) + torch.tensor([0.0, 0.0, -3.0])
drv._global_map_dict = {"positions": pos.to(dev), "num": F_ALL}
c2ws = torch.eye(4).repeat(K, 1, 1)
c2ws[:, :3, 3] = torch.rand(K, 3, generator=g) - 0.5
rgbds = torch.rand(K, 480, 640, 4, generator=g)
rgbds[..., 3] = rgbds[..., 3] * 3.0 + 2.5
drv._c_c2w_tensor, drv._nc_rgbd_tensor, drv._frame_cid_to_ncid = c2ws.to(dev), rgbds.to(dev), torch.arange(K, device=dev)
cur = torch.arange(20, device=dev)


def once():
    optim.adam_step(params, vm, state, ids, 1e-3, 1e-15, 1e-5)
    t = targets.sample_target_mv(drv, cur)
    o = targets.get_observed_fields(drv, drv._nc_rgbd_tensor[3], drv._c_c2w_tensor[3])
    torch.cuda.synchronize()
    return t, o


for _ in range(2):
    once()
t, o = once()
elems = sum(v[0].numel() for v in params.values()) * F_ACT
print(f"adam elements {elems} ({elems * 36} algorithmic bytes); target fields {len(t.field_ids)}; observed {len(o)}")
