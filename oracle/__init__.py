"""CPU oracle for the articulated-Gaussian-splat render path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  ``manus_b200`` never imports this package.

Parity status
-------------
* pose path (P1-P4: LBS, covariance, SH->RGB, activations): PINNED against the
  reference's own Python (``/root/reference/src``) through the golden vectors in
  ``tests/golden/pose_golden_*.npz`` made by ``tests/golden/make_golden_pose.py``.
* image losses (``loss_ref.py``) and the skin-weight lookup (``skin_ref.py``): PINNED against the reference's own
  ``loss_utils`` / ``skinning_weights_from_voxel_grid`` through ``tests/golden/loss_golden.npz`` / ``skin_golden.npz``
  (``tests/golden/make_golden_loss.py``, ``make_golden_skin.py``).
* contact distance (``knn_ref.contact_dist``): the reference loop is a taichi kernel (cannot run here); cross-checked with a k-d
  tree and pinned through the reference's pure-torch ``get_contact_map`` (tests/golden/contact_golden.npz).
* rasterizer (R2/R3) and ``distCUDA2`` (K1): **parity unpinned** -- the Inria
  ``diff-gaussian-rasterization`` / ``simple-knn`` sources are not present in
  ``/root/reference`` (cloned un-pinned at install time, ``setup_env.sh:4-13``);
  the restatement follows their published algorithm as recorded in SURVEY.md
  Appendix A / B and is cross-checked by an independent autograd restatement and
  analytic known-answer tests only.
"""
