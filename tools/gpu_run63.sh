#!/bin/bash
timeout 300 python -m pytest tests/test_shard_gpu.py -m gpu -q -x -k "two_rank" 2>&1 | tail -4
