"""point_diffusion_refinement_b200: the data-parallel hot path of ZhaoyangLyu/Point_Diffusion_Refinement
(1000-step DDPM reverse sampling over the dual-path PointNet++ denoiser, plus the Chamfer / EMD
evaluation kernels) rebuilt for NVIDIA B200 (sm_100a).

Layout: ``csrc/`` hand-written CUDA behind the C ABI of ``include/pdr_b200.h`` -> ``libpdr_b200.so``;
``_lib`` ctypes binding; ``_ext`` / ``knn`` / ``emd_cuda`` drop-ins for the reference's native modules;
``pointnet2_utils`` / ``pointnet2_modules`` / ``attention`` / ``pointnet2_ssg_sem`` /
``pointnet2_with_pcld_condition`` / ``pnet`` / ``chamfer_loss_new`` / ``emd`` / ``util`` /
``util_fastdpmv2`` host-side mirrors of the reference's operator API; ``dropin`` to run the unmodified
reference Python on these kernels; ``dist`` for the one-process-per-GPU shard + final gather.

There is no CPU path: importing is cheap, calling any op without the CUDA library or a GPU raises.
"""
__version__ = "0.1.0"
