#!/bin/sh
# Builds the plain-C part of the oracle (test infrastructure).  Output: oracle/_build/liblsap_oracle.so
set -e
here="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$here/_build"
gcc -O2 -fPIC -shared -fno-fast-math -ffp-contract=off -o "$here/_build/liblsap_oracle.so" "$here/lsap.c" -lm
