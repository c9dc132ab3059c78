#!/bin/bash
# The bench frame with the offsets already in HBM, rendered in chunks of N Mi samples (two chunks in flight): what the
# chunking itself costs, without the upload.
for c in ${CHUNKS:-0 8 16 32}; do
  echo -n "chunk $c Mi: "
  python scripts/profile_frame.py --frames 6 --no-profile --chunk $((c << 20)) 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('total %.2f ms, %d chunks' % (d['ms_total'], d['chunks']))"
done
