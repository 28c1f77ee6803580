#!/bin/bash
echo "u8 default"; timeout 120 python scripts/exp_graph_latency.py u8 2>&1 | tail -1
echo "u8 inline prepare"; LSD_NO_PREP_THREAD=1 timeout 120 python scripts/exp_graph_latency.py u8 2>&1 | tail -1
echo "u8 1 pack thread"; LSD_PACK_THREADS=1 timeout 120 python scripts/exp_graph_latency.py u8 2>&1 | tail -1
echo "u8 inline + 1 pack thread"; LSD_NO_PREP_THREAD=1 LSD_PACK_THREADS=1 timeout 120 python scripts/exp_graph_latency.py u8 2>&1 | tail -1
