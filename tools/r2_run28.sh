#!/bin/bash
for ips in 32 64 128 256; do
  echo "items per SM $ips"
  FC_PRUNE_ITEMS_PER_SM=$ips FC_PRUNE_TRACE=1 python tools/run_c4.py 200000 2>&1 | grep -E "fc_prune: total|pass k=(2|1) " | tail -3
done
