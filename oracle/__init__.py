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
  * SMPL forward (LBS): **parity unpinned** -- the arithmetic lives in third-party
    smplx==0.1.28 wrapped by pare==0.1 (environment.yml:118,167), neither vendored nor
    installable offline, and the reference tree holds no golden vertices.  Mitigation: two
    independent restatements (batched torch `smpl_oracle.py` following the published smplx
    lbs.py, per-body NumPy/cv2 `smpl_webuser_oracle.py` following the in-tree
    models/smpl_webuser/) must agree, plus analytic invariants.
"""
