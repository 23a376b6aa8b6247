# usage: ab_lib.sh <people> lib1.so lib2.so ...  -- bench with alternative builds of the library
n=$1; shift
cp june_b200/libjune_b200.so /tmp/orig.so
for l in "$@"; do cp june_b200/$l june_b200/libjune_b200.so; bash tools/bench_line.sh $n LIB=$l; done
cp /tmp/orig.so june_b200/libjune_b200.so
