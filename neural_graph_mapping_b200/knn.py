"""kNN (use_vmap=False) branch of the field set -- ngm/models.py:347-405.  Placeholder until the
CUDA kernels land."""


def fieldset_forward_knn(model, query_points, field_positions, field_orientations, field_ids, field_radius):
    raise NotImplementedError("kNN field-set path (models.py:347-405) is not built in this revision")


def render_rays_knn(driver, ijs, c2ws, camera, field_ids, near, far, gt, overwrite, jitter):
    raise NotImplementedError("kNN render path (use_vmap=False) is not built in this revision")
