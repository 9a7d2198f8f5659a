#!/bin/bash
run() { python bench.py --no-cpu-baseline --no-configs --no-train-step --no-e2e --steps 64 2>/dev/null | tail -1 > gpurun_out/c45.json; python -c "
import json; d=json.load(open('gpurun_out/c45.json')); print('$1', d['value'], sorted(d['batches']['device_ms'])[:3])"; }
run base
GSB_BWD_CTAS=2 run bwd2
GSB_BWD_CTAS=1 run bwd1
GSB_FWD_PAD_KB=60 run fwdpad60_3ctas
GSB_FWD_PAD_KB=100 run fwdpad100_2ctas
GSB_BWD_CTAS=2 GSB_FWD_PAD_KB=60 run bwd2_fwd3
GSB_BWD_CTAS=2 GSB_FWD_PAD_KB=100 run bwd2_fwd2
