#!/bin/bash
# Build an A/B variant of the library with extra nvcc defines:  tools/build_variant.sh NAME -DFOO=1 ...
# -> gpurun_variants/libcdetr_NAME.so; select it with CDETR_LIB_PATH (counting_detr_b200/_lib.py).
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/gpurun_variants; mkdir -p $OUT/obj_$NAME
FLAGS=$(python - <<PY
import sys; sys.path.insert(0, "$ROOT")
import __graft_entry__ as g; print(" ".join(g.NVCC_FLAGS))
PY
)
for f in $ROOT/counting_detr_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc $FLAGS "$@" -c $f -o $OUT/obj_$NAME/$b.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $OUT/libcdetr_$NAME.so $OUT/obj_$NAME/*.o -lcudart
rm -rf $OUT/obj_$NAME
ls -la $OUT/libcdetr_$NAME.so
