#!/bin/bash
# e2e frame (host offsets streamed in) for a few streaming chunk sizes: RAYHS_B200_STREAM_CHUNK_MI x RAYHS_B200_STREAM_FIRST_MI
for c in ${CHUNKS:-8 16 24 32 48}; do
  for f in ${FIRSTS:-4}; do
    echo -n "chunk $c Mi first $f Mi tail ${RAYHS_B200_TAIL:-1}: "
    RAYHS_B200_STREAM_CHUNK_MI=$c RAYHS_B200_STREAM_FIRST_MI=$f python scripts/e2e_frame.py --frames 10 2>&1 | tail -1
  done
done
