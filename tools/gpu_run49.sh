#!/bin/bash
timeout 300 python tools/e2e_probe.py 8192
