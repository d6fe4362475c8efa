"""Load ``tests/golden/*.npz`` (made by ``oracle/make_golden.py`` from the unmodified
reference) and turn the stored reference kwargs into oracle specs."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_ENC = {
    "PositionalEncodingNeRF": "nerf",
    "PositionalEncodingFourier": "fourier",
    "TriplaneEncoding": "triplane",
    "PermutohedralEncoding": "permuto",
}


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta_json"}
    return meta, arrays


def encoding_kind(field_kwargs):
    return _ENC[field_kwargs["encoding_type"].split(".")[-1]]


def field_spec(field_kwargs):
    from oracle import restatement as R

    return R.FieldSpec(
        encoding=encoding_kind(field_kwargs),
        encoding_kwargs=dict(field_kwargs["encoding_kwargs"]),
        num_layers=field_kwargs["num_layers"],
        dim_out=field_kwargs["dim_out"],
        dim_mlp_out=field_kwargs.get("dim_mlp_out"),
        skip_mode=field_kwargs.get("skip_mode", "no"),
    )


def render_spec(meta):
    from oracle import restatement as R

    cfg = meta["config"]
    mk = cfg["model_kwargs"]
    rdg = cfg.get("range_depth_guided")
    return R.RenderSpec(
        num_samples=meta["num_samples"],
        num_samples_depth_guided=meta.get("num_samples_depth_guided", 0),
        range_depth_guided=cfg["truncation_distance"] if rdg is None else rdg,
        near_distance=meta.get("near_distance", cfg["near_distance"]),
        far_distance=meta.get("far_distance", cfg["far_distance"]),
        truncation_distance=cfg["truncation_distance"],
        freespace_weight=cfg["freespace_weight"],
        tsdf_weight=cfg["tsdf_weight"],
        geometry_mode=cfg["geometry_mode"],
        geometry_factor=cfg["geometry_factor"],
        color_factor=cfg["color_factor"],
        field_radius=mk["field_radius"],
        scale_mode=mk["scale_mode"],
        num_knn=mk["num_knn"],
        distance_factor=mk["distance_factor"],
        outside_value=mk["outside_value"],
        block_size=cfg["block_size"],
        pixel_block_size=cfg["pixel_block_size"],
    )


def camera_spec(cam):
    from oracle import restatement as R

    return R.CameraSpec(**cam)


def params(arrays, prefix="param:"):
    return {k[len(prefix):]: v for k, v in arrays.items() if k.startswith(prefix)}


VMAP_CASES = ["c1_vmap_256x32", "vmap_guided_nrgbd", "vmap_behind_camera", "vmap_neus",
              "vmap_density", "vmap_occupancy", "c2_vmap_w128_s64"]
KNN_CASES = ["knn_render", "knn_render_w128"]


def replay_training(meta, a, iteration_fn, grow_fn):
    """Drive ``iteration_fn(it, inputs)`` / ``grow_fn(rows)`` through the fixture's schedule."""
    for st in meta["steps"]:
        if "grow" in st:
            it = st["before_iteration"]
            grow_fn({k[len(f"grow{it}:param:"):]: v for k, v in a.items() if k.startswith(f"grow{it}:param:")})
        else:
            it = st["iteration"]
            iteration_fn(it, {k[len(f"it{it}:"):]: v for k, v in a.items() if k.startswith(f"it{it}:")})


def rigid_from_quaternion(q, t):
    """(..., 4, 4) camera/world transform from a real-first unit quaternion and a translation."""
    w, x, y, z = q.unbind(-1)
    rot = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                       2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                       2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).view(*q.shape[:-1], 3, 3)
    T = torch.eye(4).repeat(*q.shape[:-1], 1, 1)
    T[..., :3, :3] = rot
    T[..., :3, 3] = t
    return T


def quaternion_multiply(a, b):
    """Real-first Hamilton product (pytorch3d.transforms.quaternion_multiply's convention up to sign)."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def loop_closure_update(positions, orientations, kf_ids, delta_q, delta_t):
    """What NeuralGraphMap._update_field_poses does to the fields when keyframe k's pose changes from P_k to
    D_k P_k (ngm/run_mapping.py:846-884, 937-952: absolute -> relative to the anchoring keyframe with the old poses,
    relative -> absolute with the new ones, i.e. X' = D_k X): positions D_k p, orientations q(D_k) (x) q."""
    D = rigid_from_quaternion(delta_q, delta_t)[kf_ids]
    new_pos = torch.einsum("fdk,fk->fd", D[:, :3, :3], positions) + D[:, :3, 3]
    return new_pos, quaternion_multiply(delta_q[kf_ids], orientations), D
