#!/usr/bin/env bash
# dev: time the default library and every tools/variants/*.so with tools/kx_time.py
cd "$(dirname "$0")/.."
echo "== default"; python tools/kx_time.py
for so in tools/variants/*.so; do
  echo "== $so"; VPDQ_B200_LIB=$PWD/$so timeout 120 python tools/kx_time.py
done
