#!/bin/bash
# Per-warp begin / end times of the pooled shadow kernel per pass (diagnostic build variants/wt, -DRH_WARP_TIMES) on the
# bench frame and on shard 0 of 8; with COUNT=1 also the distribution of the walks' lengths (instrumented kernels).
RAYHS_B200_LIB=variants/wt/librayhs_b200.so RAYHS_B200_DEBUG=1 python scripts/shard_frame.py --frames 1 ${COUNT:+--count} 2>&1 | grep -v "^{" | grep -A2 "pooled pass" | tail -16 | cut -c1-600
