#!/bin/bash
# Last check of a build: GPU tests, smoke(), one timed bench frame.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python scripts/profile_frame.py --frames 4 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('frame ms %.2f trace %.2f shadow %.2f' % (d['ms_total'], d['ms_trace'], d['ms_shadow']))"
