#!/bin/bash
# A/B timing of two builds of the library on the SAME GPU box (boxes differ by several percent): alternates
#   DTRAJ_LIB=<A> / DTRAJ_LIB=<B>  python tools/profile_layers.py <args...>
# and prints the per-model totals and the named layer of every run.     usage: tools/ab.sh <libA> <libB> <layer regex> <profile_layers args...>
A=$1; B=$2; PAT=$3; shift 3
for rep in 1 2 3; do
  for L in "$A" "$B"; do
    echo "## $(basename $L) run $rep"
    DTRAJ_LIB=$L python tools/profile_layers.py "$@" | grep -E "==|$PAT"
  done
done
