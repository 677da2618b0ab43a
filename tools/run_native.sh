#!/bin/bash
# Runs every case of a native test binary, each in its own process under a timeout.
# usage: tools/run_native.sh <binary> [logfile]
BIN=$1; LOG=${2:-/dev/stdout}
N=$($BIN)
for i in $(seq 0 $((N-1))); do
  timeout 120 $BIN $i >> $LOG 2>&1 || echo "CASE $i exit=$?" >> $LOG
done
