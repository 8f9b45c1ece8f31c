#!/bin/bash
cd "$(dirname "$0")/.."
GB_DEFINES="${1:-}" python gordo_components_b200/csrc/build.py > /dev/null || exit 1
GB_DEFINES="${1:-}" timeout -k 10 600 python scratch/lstm_trace_digest.py 2>&1 | tail -30
