"""Install the B200 kernels underneath the UNMODIFIED reference Python.

    import point_diffusion_refinement_b200.dropin as dropin
    dropin.install()                      # before `import pointnet2_ops` / `import pointnet2...`
    sys.path[:0] = [REF + "/pointnet2_ops_lib", REF, REF + "/pointnet2"]
    from pointnet2.models.pointnet2_with_pcld_condition import PointNet2CloudCondition   # reference class

After ``install()`` the reference's own modules resolve
  * ``pointnet2_ops._ext``                     -> point_diffusion_refinement_b200._ext      (9 functions)
  * ``pytorch3d.ops.knn`` / ``pytorch3d.ops``  -> point_diffusion_refinement_b200.knn
  * ``pytorch3d.structures.pointclouds``       -> stub with ``Pointclouds`` (isinstance check only)
  * ``emd_cuda``                               -> point_diffusion_refinement_b200.emd_cuda
(import sites: pointnet2_ops/pointnet2_utils.py:7-10, pointnet2/chamfer_loss_new.py:6-7, pointnet2/emd.py:2).
"""
import contextlib
import sys
import types


def install(ext=None, knn_module=None, emd_module=None):
    """Register the shims in ``sys.modules``.  The keyword arguments exist for the tests, which bind the
    reference Python to the CPU oracle instead; product code calls ``install()`` with no arguments."""
    from . import _ext as _pdr_ext, emd_cuda as _pdr_emd, knn as _pdr_knn
    ext = ext or _pdr_ext
    knn_module = knn_module or _pdr_knn
    emd_module = emd_module or _pdr_emd
    sys.modules["pointnet2_ops._ext"] = ext
    p3d = types.ModuleType("pytorch3d")
    ops = types.ModuleType("pytorch3d.ops")
    structures = types.ModuleType("pytorch3d.structures")
    pcl = types.ModuleType("pytorch3d.structures.pointclouds")
    pcl.Pointclouds = getattr(knn_module, "Pointclouds", type("Pointclouds", (), {}))
    ops.knn = knn_module
    ops.knn_points = knn_module.knn_points
    ops.knn_gather = knn_module.knn_gather
    structures.pointclouds = pcl
    structures.Pointclouds = pcl.Pointclouds
    p3d.ops, p3d.structures = ops, structures
    sys.modules.update({"pytorch3d": p3d, "pytorch3d.ops": ops, "pytorch3d.ops.knn": knn_module,
                        "pytorch3d.structures": structures, "pytorch3d.structures.pointclouds": pcl,
                        "emd_cuda": emd_module})


def uninstall():
    for name in ("pointnet2_ops._ext", "pytorch3d", "pytorch3d.ops", "pytorch3d.ops.knn", "pytorch3d.structures",
                 "pytorch3d.structures.pointclouds", "emd_cuda"):
        sys.modules.pop(name, None)


@contextlib.contextmanager
def reference_torch_version(version="1.7.1"):
    """``completion_eval.py:14-32`` refuses to import unless torch.__version__ is '1.7.1' or '1.4.0'.
    Wrap the import of the reference drivers in this context."""
    import torch
    real = torch.__version__
    torch.__version__ = version
    try:
        yield
    finally:
        torch.__version__ = real
