#!/usr/bin/env bash
# Build the host simulation of the device headers (test tool, see hostsim.cpp).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
g++ -O2 -std=c++17 -ffp-contract=off -mfma -msse4.1 -Wall -Wno-unused-function -Wno-maybe-uninitialized -o "$HERE/hostsim" "$HERE/hostsim.cpp"
