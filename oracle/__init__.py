"""CPU oracle for the W-HMR body-model hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package; the product path (`w-hmr_b200/`) never does and fails loudly
when its CUDA library is missing.

Pinning status (repeated in DESIGN.md):
  * projection / perspective_projection / convert_pare_to_full_img_cam / batch_rodrigues
    (quaternion variant) / rot6d / gram-schmidt / rotmat->axis-angle / MAF_Extractor.sampling
    / Procrustes: PINNED against outputs of the reference's own code imported in the
    authoring container (tests/golden/make_golden.py -> tests/golden/*.npz) and against the
    reference's vendored known-answer tests (Procrustes scale+translate round trip,
    identity Rodrigues).
  * SMPL forward (LBS), vertices and the 24 posed chain joints: PINNED against outputs of the reference's
    own in-tree SMPL code -- models/smpl_webuser/{serialization,posemapper,verts,lbs}.py, the original
    SMPL loader the batched smplx/pare wrapper is a twin of -- EXECUTED in the authoring container
    (tests/golden/make_golden_smpl.py: numpy `xp` branch + cv2.Rodrigues, a stand-in for the absent
    `chumpy` import; -> tests/golden/smpl_webuser_outputs.npz, 16 bodies incl. the reference's 8 real
    fixture poses, theta=0 and a joint at ~pi) on the synthetic model.  fp64 oracle: 2e-7 m of the goldens.
  * What stays a restatement of the published smplx==0.1.28 / pare==0.1 code (environment.yml:118,167;
    third-party, neither vendored nor installable offline): the batch_rodrigues epsilon variant
    (angle = ||theta + 1e-8||, differs from cv2.Rodrigues by O(1e-8)), the batched tensor layout, and the
    49-joint assembly (VertexJointSelector + J_regressor_extra + joint_map) -- anchored on the in-tree
    commented twin models/smpl.py:61-83, its call sites, and cross-checked by `smpl_webuser_oracle.py`
    plus analytic invariants.
"""
