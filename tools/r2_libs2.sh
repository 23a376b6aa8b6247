#!/bin/bash
# usage: r2_libs2.sh "<people list>" lib1.so lib2.so ... -- step timeline with alternative builds of the library
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
PL=$1; shift
cp june_b200/libjune_b200.so /tmp/main.so
for L in "$@"; do
  cp june_b200/$L june_b200/libjune_b200.so
  for P in $PL; do
    echo "== $L people $P"
    JB_STAMPS=1 timeout 600 python tools/stamps.py $P 2>&1 | grep -v "last block" | tail -2
  done
done
cp /tmp/main.so june_b200/libjune_b200.so
