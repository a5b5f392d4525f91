#!/bin/bash
# Rebuild libb2deflate.so + the oracle, then run a command on the GPU box.  usage: tools/gpu.sh [--timeout S] -- '<cmd>'
set -e
cd /root/repo
python /root/repo/deflate-library-java_b200/build.py > /tmp/build.log 2>&1 || { cat /tmp/build.log; exit 1; }
make -C /root/repo/oracle -s
exec /usr/local/graft/bin/gpurun "$@"
