#!/bin/bash
# AddressSanitizer + UBSan over the host emulation of the kernel source (the CPU stand-in for compute-sanitizer memcheck when no
# GPU is at hand): rebuilds one emulation flavor instrumented, runs the host tests that drive it, restores the plain library.
#   bash tests/asan_hostsim.sh            # regular flavor: Fetch host env tests incl. the NaN-injection recovery test
#   bash tests/asan_hostsim.sh kitchen    # -DB200_KITCHEN: two-level broad phase, joint equalities, condim 6 (soak + env parity)
flavor=${1:-plain}
cd "$(dirname "$0")/.."
if [ "$flavor" = kitchen ]; then lib=tests/hostsim/libhostsim_kitchen.so; defs="-DB200_KITCHEN"; sel="tests/test_soak_host.py tests/test_kitchen_host.py -k kitchen"
else lib=tests/hostsim/libhostsim.so; defs=""; sel="tests/test_host_env.py"; fi
cp $lib /tmp/asan_backup.so
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -fPIC -shared -Wno-unused-function $defs -o $lib tests/hostsim/hostsim.cpp
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" timeout 900 python -m pytest $sel -x -q 2>&1 | tail -12
cp /tmp/asan_backup.so $lib
