"""Constants the body-model hot path depends on.

Every value here is part of the op contract of the reference (yw0208/W-HMR) and is
cited to the file:line it comes from; see SURVEY.md Appendix A.
"""

# core/constants.py:4 -- focal length used by utils/geometry.py:projection
FOCAL_LENGTH = 1000.0
# configs/pymaf_config.yaml:83-85 -- cfg.IMG_RES.{WIDTH,HEIGHT} read inside projection()
IMG_RES_WIDTH = 256
IMG_RES_HEIGHT = 256
# configs/pymaf_config.yaml:36 -- cfg.MODEL.PyMAF.MLP_DIM (MAF_Extractor.reduce_dim)
MLP_DIM = (256, 128, 64, 32)
# configs/pymaf_config.yaml:37
N_ITER = 3

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_BASIS = 207  # 23 * 9, models/smpl_webuser/posemapper.py:36-43 (lrotmin)

# SMPL kinematic tree (kintree_table[0]); parents[0] = -1.  smplx==0.1.28 stores
# exactly this table for SMPL; twin: models/smpl_webuser/lbs.py:30-31.
SMPL_PARENTS = (-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21)

# models/smpl.py:14-32 JOINT_MAP and :33-51 JOINT_NAMES  ->  49 indices into the
# 54-joint set [24 chain joints | 21 selected vertices | 9 J_regressor_extra joints].
JOINT_MAP = {
    'OP Nose': 24, 'OP Neck': 12, 'OP RShoulder': 17,
    'OP RElbow': 19, 'OP RWrist': 21, 'OP LShoulder': 16,
    'OP LElbow': 18, 'OP LWrist': 20, 'OP MidHip': 0,
    'OP RHip': 2, 'OP RKnee': 5, 'OP RAnkle': 8,
    'OP LHip': 1, 'OP LKnee': 4, 'OP LAnkle': 7,
    'OP REye': 25, 'OP LEye': 26, 'OP REar': 27,
    'OP LEar': 28, 'OP LBigToe': 29, 'OP LSmallToe': 30,
    'OP LHeel': 31, 'OP RBigToe': 32, 'OP RSmallToe': 33, 'OP RHeel': 34,
    'Right Ankle': 8, 'Right Knee': 5, 'Right Hip': 45,
    'Left Hip': 46, 'Left Knee': 4, 'Left Ankle': 7,
    'Right Wrist': 21, 'Right Elbow': 19, 'Right Shoulder': 17,
    'Left Shoulder': 16, 'Left Elbow': 18, 'Left Wrist': 20,
    'Neck (LSP)': 47, 'Top of Head (LSP)': 48,
    'Pelvis (MPII)': 49, 'Thorax (MPII)': 50,
    'Spine (H36M)': 51, 'Jaw (H36M)': 52,
    'Head (H36M)': 53, 'Nose': 24, 'Left Eye': 26,
    'Right Eye': 25, 'Left Ear': 28, 'Right Ear': 27,
}
JOINT_NAMES = (
    'OP Nose', 'OP Neck', 'OP RShoulder', 'OP RElbow', 'OP RWrist', 'OP LShoulder',
    'OP LElbow', 'OP LWrist', 'OP MidHip', 'OP RHip', 'OP RKnee', 'OP RAnkle',
    'OP LHip', 'OP LKnee', 'OP LAnkle', 'OP REye', 'OP LEye', 'OP REar',
    'OP LEar', 'OP LBigToe', 'OP LSmallToe', 'OP LHeel', 'OP RBigToe', 'OP RSmallToe',
    'OP RHeel', 'Right Ankle', 'Right Knee', 'Right Hip', 'Left Hip', 'Left Knee',
    'Left Ankle', 'Right Wrist', 'Right Elbow', 'Right Shoulder', 'Left Shoulder',
    'Left Elbow', 'Left Wrist', 'Neck (LSP)', 'Top of Head (LSP)', 'Pelvis (MPII)',
    'Thorax (MPII)', 'Spine (H36M)', 'Jaw (H36M)', 'Head (H36M)', 'Nose', 'Left Eye',
    'Right Eye', 'Left Ear', 'Right Ear',
)
JOINT_MAP_49 = tuple(JOINT_MAP[n] for n in JOINT_NAMES)
assert len(JOINT_MAP_49) == 49 and max(JOINT_MAP_49) == 53

# models/smpl.py:57-58
H36M_TO_J17 = (6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9)
H36M_TO_J14 = H36M_TO_J17[:14]

# smplx==0.1.28 vertex_ids['smplh'] (models/whmr.py:60 VertexJointSelector(vertex_ids['smplh'])).
# The table is NOT in the reference tree (third-party); values below are the published
# smplx ones.  Selector order: 5 face, 6 feet, then left/right x (thumb,index,middle,ring,pinky).
# Every op takes the index list as an input, so parity never depends on these numbers.
SMPLH_VERTEX_IDS = {
    'nose': 332, 'reye': 6260, 'leye': 2800, 'rear': 4071, 'lear': 583,
    'rthumb': 6191, 'rindex': 5782, 'rmiddle': 5905, 'rring': 6016, 'rpinky': 6133,
    'lthumb': 2746, 'lindex': 2319, 'lmiddle': 2445, 'lring': 2556, 'lpinky': 2673,
    'LBigToe': 3216, 'LSmallToe': 3226, 'LHeel': 3387,
    'RBigToe': 6617, 'RSmallToe': 6624, 'RHeel': 6787,
}


def vertex_joint_selector_ids(vertex_ids=None):
    """Index list in smplx VertexJointSelector order (21 entries for SMPL+H ids)."""
    v = SMPLH_VERTEX_IDS if vertex_ids is None else vertex_ids
    ids = [v['nose'], v['reye'], v['leye'], v['rear'], v['lear']]
    ids += [v['LBigToe'], v['LSmallToe'], v['LHeel'], v['RBigToe'], v['RSmallToe'], v['RHeel']]
    for hand in ('l', 'r'):
        for tip in ('thumb', 'index', 'middle', 'ring', 'pinky'):
            ids.append(v[hand + tip])
    return ids
