#!/bin/bash
# e2e throughput of b200_compress_batch over the knobs that trade latency for parallelism (run on a GPU box).
# usage: tools/sweep_e2e.sh "MEGABATCH:WORKERS:SUBSEQ ..." [n_images]
N=${2:-512}
for cfg in $1; do
  IFS=: read mb w ss <<< "$cfg"
  echo "== megabatch=$mb workers=$w subseq=$ss"
  B200_MEGABATCH=$mb B200_GROUP_WORKERS=$w B200_DEC_SUBSEQ=$ss REPS=3 timeout 120 python tools/throughput.py 16 $N 2>&1 | grep "C-ABI only" | tail -2
done
