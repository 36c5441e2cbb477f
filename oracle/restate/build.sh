#!/bin/bash
# Builds the plain-C restatement of the traversal / sampler core (test infrastructure, see rt_restate.c).
set -e
cd "$(dirname "$0")"
mkdir -p ../_ref/restate
if [ ../_ref/restate/rt_restate -nt rt_restate.c ] && [ ../_ref/restate/rt_restate -nt ../../include/rtb200.h ]; then exit 0; fi
gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -mfma -msse4.1 -Wall -Wextra -o ../_ref/restate/rt_restate rt_restate.c -lm
