#!/bin/bash
# Build everything that travels to the GPU box, then run `gpurun`: tools/grun.sh [gpurun options] -- command
set -e
cd "$(dirname "$0")/.."
make -s -C rtl-wmbus_b200/csrc
make -s -C oracle oracle
[ -e /root/reference/rtl_wmbus.c ] && make -s -C oracle ref
exec /usr/local/graft/bin/gpurun "$@"
