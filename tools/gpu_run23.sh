#!/bin/bash
set -x
python tools/latency_probe.py 2>&1 | tail -20
