import sys, math
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import test_gpu_mapping_loop as T
for prec in ("fp32", "fp16"):
    loop, cam, stream = T._loop(prec)
    loop._update_slam_state(0)
    print(prec, "fields", loop._num_fields, "current", loop._current_field_ids.tolist())
    for it in range(3):
        tgt = None
        out = loop._optimization_iteration()
        tgt = loop._target
        print(" it", it, {k: float(v) for k, v in out.items()})
        print("   target fields", tgt.field_ids.tolist(), "depth_mask", int(tgt.depth_mask.sum()), "term_mask", int(tgt.term_mask.sum()),
              "gt nan", bool(tgt.gt_distances.isnan().any()), "gt range", float(tgt.gt_distances.min()), float(tgt.gt_distances.max()),
              "near/far", float(tgt.near_distances.min()), float(tgt.far_distances.max()))
        bad = [k for k, v in loop._model.all_fields_params.items() if not torch.isfinite(v).all()]
        print("   non-finite params:", bad)
