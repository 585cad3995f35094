#!/usr/bin/env python
"""Index tables of the reference, extracted by IMPORTING the reference's own module (authoring container only:
needs /root/reference):  python tests/golden/make_golden_tables.py  ->  reference_tables.npz

  models/smpl.py:14-58   JOINT_MAP + JOINT_NAMES  -> the 49-entry index list SMPL.forward applies to the 54-joint
                         set (models/smpl.py:66-76), H36M_TO_J17, H36M_TO_J14
  core/constants.py      FOCAL_LENGTH (parsed: the module imports core.cfgs -> yacs, which is absent)

`oracle/smpl_oracle.py` reads this file (not the product package's constants), and tests/test_oracle_cpu.py asserts
that `whmr_b200.constants` equals it entry by entry."""
import importlib.util
import os
import re

import numpy as np

REF = os.environ.get('WHMR_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location('ref_models_smpl', os.path.join(REF, 'models', 'smpl.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    names = list(m.JOINT_NAMES)
    joint_map_49 = np.asarray([m.JOINT_MAP[n] for n in names], dtype=np.int64)      # models/smpl.py:66-68
    src = open(os.path.join(REF, 'core', 'constants.py')).read()
    focal = float(re.search(r'^FOCAL_LENGTH\s*=\s*([0-9.]+)', src, re.M).group(1))
    np.savez(os.path.join(HERE, 'reference_tables.npz'),
             joint_names=np.asarray(names), joint_map_49=joint_map_49,
             joint_map_keys=np.asarray(list(m.JOINT_MAP.keys())),
             joint_map_vals=np.asarray(list(m.JOINT_MAP.values()), dtype=np.int64),
             h36m_to_j17=np.asarray(m.H36M_TO_J17, dtype=np.int64), h36m_to_j14=np.asarray(m.H36M_TO_J14, dtype=np.int64),
             focal_length=np.float64(focal))
    print('joint_map_49', joint_map_49.tolist())
    print('h36m_to_j14', list(m.H36M_TO_J14), 'focal', focal)


if __name__ == '__main__':
    main()
