#!/usr/bin/env python
"""Compare two result files in the reference's exp/result format (frames matched by name):
    python tools/compare_results.py /path/to/reference/exp/result/icvl.txt exp/result/icvl_b200.txt
With the authors' checkpoint restored (python -m densereg_b200.model --dataset icvl --is_train False) and the ICVL test shards in place, this is
the row-for-row comparison against the reference's published predictions."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densereg_b200.model import compare_result_files
if len(sys.argv) != 3:
    sys.exit(__doc__)
c = compare_result_files(sys.argv[1], sys.argv[2])
c["curve"] = [list(t) for t in c["curve"]]
print(json.dumps(c, indent=1))
