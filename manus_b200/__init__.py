"""manus_b200 -- B200-native (sm_100a) articulated-Gaussian-splat render path behind MANUS's own operator boundary.

  rasterizer : GaussianRasterizationSettings / GaussianRasterizer  (drop-in for diff_gaussian_rasterization)
  knn        : distCUDA2                                           (drop-in for simple_knn._C)
  pose       : pose_gaussians / bone_transforms                    (fused LBS + covariance + SH->RGB, fwd + bwd)
  render     : render_gaussians / render_fused                     (mirror of src/utils/gaussian_utils.py:349-428)
  dist       : view-sharded data parallelism (one NCCL all-reduce of the flat per-Gaussian gradient buffer)
  cameras, synth : host-side camera matrices and the seeded synthetic scenes used by the tests and the bench

All arithmetic runs in manus_b200/lib/libmanus_b200.so (C ABI: include/manus_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
