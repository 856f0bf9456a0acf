#!/bin/bash
# TEST INFRASTRUCTURE ONLY: builds the CPU simulation of libwmbus_b200 used by the
# `-m "not gpu"` tests (see hostsim_cuda.h).  Output: tests/hostsim/_build/libwmbus_hostsim.so
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
src="$root/rtl-wmbus_b200/csrc"
mkdir -p "$here/_build"
gcc -O2 -std=gnu99 -fPIC -Wall -Wextra -I"$root/include" -c "$src/wmb_framer.c" -o "$here/_build/wmb_framer.o"
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -Wall -Wextra -Wno-unused-parameter -Wno-unused-function -Wno-unknown-pragmas -Wno-stringop-overflow \
    -DWMB_HOSTSIM -DWMB_VERSION='"wmbus-b200 hostsim (TEST ONLY)"' \
    -I"$here" -I"$root/include" -I"$src" -x c++ -c "$src/wmb_context.cu" -o "$here/_build/wmb_context.o"
g++ -shared -o "$here/_build/libwmbus_hostsim.so" "$here/_build/wmb_context.o" "$here/_build/wmb_framer.o" -lm -lpthread
echo "built $here/_build/libwmbus_hostsim.so"
# the drop-in host program against the CPU simulation (live-stream / watchdog tests without a GPU)
gcc -O2 -std=gnu99 -Wall -Wextra -I"$root/include" -o "$here/_build/rtl_wmbus_hostsim" "$src/rtl_wmbus_b200.c" \
    -L"$here/_build" -lwmbus_hostsim -Wl,-rpath,"$here/_build" -lstdc++ -lm
echo "built $here/_build/rtl_wmbus_hostsim"
