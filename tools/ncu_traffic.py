#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` raw CSV of tools/profile_smpl.py --loop-batch B:
dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by the kernel names bench.py reports.
usage: ncu_traffic.py <raw.csv> <B> <source label> [out.json] [launches of the loop pass, default 18]"""
import csv
import json
import os
import sys


def main(path, B, source, out, n_last=18):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    ri, wi = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    seq = [(r[ki], float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]) for r in rows[2:]]
    # the loop pass is the LAST n_last launches of the capture (the stand-alone head call comes first);
    # 18 on the deferred schedule (RegressorLoop.defer), 22 on the immediate one
    names = {'smpl_chain_kernel': 'chain', 'smpl_fused_tc_kernel': 'blend_skin', 'readout_reduce_kernel': 'readout',
             'project_weak_kernel': 'project_weak', 'project_full_kernel': 'project_weak_full',
             'pose_blend_tc_kernel': 'pose_blend', 'skin_tc_kernel': 'skin'}
    acc, n_sample = {}, 0
    for name, b in seq[-n_last:]:
        key = None
        if 'sample_bilinear' in name:
            key = 'sample_l%d' % n_sample
            n_sample += 1
        else:
            for k, v in names.items():
                if k in name:
                    key = v
        if key:
            a = acc.setdefault(key, [0.0, 0])
            a[0] += b
            a[1] += 1
    res = {k: {'dram_bytes': v[0] / v[1], 'launches': v[1], 'source': source} for k, v in acc.items()}
    data = json.load(open(out)) if os.path.exists(out) else {}
    data['B%d' % B] = res
    json.dump(data, open(out, 'w'), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3],
         sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'traffic.json'),
         int(sys.argv[5]) if len(sys.argv) > 5 else 18)
