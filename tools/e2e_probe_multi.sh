#!/bin/bash
# N independent e2e probes, one per GPU, sharing the host (what N bench ranks do to it): tools/e2e_probe_multi.sh N CHANNELS SHARES
N=$1; C=$2; SH=$3
export FRA_HOST_THREADS=$(( $(nproc) / N ))
for i in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$i python tools/e2e_probe.py --channels $C --steps 12 --shares $SH > /tmp/probe_$i.txt 2>&1 &
done
wait
for i in $(seq 0 $((N-1))); do echo "== gpu $i"; grep -v trace /tmp/probe_$i.txt; done
