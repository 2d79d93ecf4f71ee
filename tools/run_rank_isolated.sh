#!/bin/bash
# test helper: every rank sees only its own GPU (CUDA_VISIBLE_DEVICES = LOCAL_RANK), the situation in which peer
# mapping fails and bench.py must fall back to the NCCL exchange
export CUDA_VISIBLE_DEVICES=$LOCAL_RANK
exec python "$@"
