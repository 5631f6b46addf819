#!/usr/bin/env python
"""Record the DRAM traffic of a kernel from an `ncu --set full` report in profiles/roofline_traffic.json, together with
a hash of the kernel's sources at capture time, so that bench.py's `roofline.traffic` says whether the number belongs to
the code that is being timed.

    python tools/traffic_from_ncu.py gpurun_out/r02_phot_final.ncu-rep phot_fused_tc
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# sources that define the traffic of each recorded kernel
SOURCES = {'phot_fused_tc': ['phot_mlp_tc.cu', 'phot_mlp_tc_body.cuh', 'mist_interp.cuh', 'tc_ptx.cuh', 'common.cuh'],
           'phot_tc': ['phot_mlp_tc.cu', 'phot_mlp_tc_body.cuh', 'tc_ptx.cuh', 'common.cuh']}
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def sources_sha256(key):
    h = hashlib.sha256()
    for f in SOURCES[key]:
        h.update(open(os.path.join(ROOT, 'minesweeper_b200', 'csrc', f), 'rb').read())
    return h.hexdigest()


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit, val = rows[0], rows[1], rows[2]                       # first profiled launch
    get = lambda k: float(val[hdr.index(k)].replace(',', '')) * UNIT[unit[hdr.index(k)]]
    rd, wr = get('dram__bytes_read.sum'), get('dram__bytes_write.sum')
    path = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = int(rd + wr)
    d[key + '_capture'] = dict(report=os.path.basename(rep), kernel=val[hdr.index('Kernel Name')],
                               dram_read_bytes=int(rd), dram_write_bytes=int(wr),
                               gpu_time_ms_under_ncu=float(val[hdr.index('gpu__time_duration.sum')]),
                               sources=SOURCES[key], sources_sha256=sources_sha256(key))
    json.dump(d, open(path, 'w'), indent=1)
    print(json.dumps({key: d[key], 'capture': d[key + '_capture']}))


if __name__ == '__main__':
    main()
