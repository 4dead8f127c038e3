#!/bin/bash
# A/B of kernel variants in ONE gpurun call. Here (no GPU): build each git ref's b200mj.cu into its own library,
#   tools/ab_variants.sh build main r2/reuse-pos
# then on the GPU box:
#   gpurun -- 'bash tools/ab_variants.sh run main r2_reuse-pos'
# which runs the parity tests of the checked-out tree once and `bench.py --no-cpu` per library, twice, interleaved.
# Only the CUDA source differs between variants; host-facade changes of a branch need that branch checked out.
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
if [ "$mode" = build ]; then
  mkdir -p gpurun_out/variants
  for ref in "$@"; do
    name=${ref//\//_}
    d=gpurun_out/variants/$name
    rm -rf $d; mkdir -p $d/dm_control_b200/csrc $d/include
    git show "$ref:dm_control_b200/csrc/b200mj.cu" > $d/dm_control_b200/csrc/b200mj.cu
    for h in b200mj.h b200mj_model_fields.h; do git show "$ref:include/$h" > $d/include/$h; done
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --shared -Xcompiler -fPIC \
        -o dm_control_b200/csrc/libb200mj_$name.so $d/dm_control_b200/csrc/b200mj.cu
    echo "built dm_control_b200/csrc/libb200mj_$name.so from $ref"
  done
else
  for rep in 1 2; do
    for name in "$@"; do
      echo -n "$name rep $rep: "
      B200MJ_SO=$PWD/dm_control_b200/csrc/libb200mj_$name.so python bench.py --no-cpu --steps 60 --warmup 20 2>&1 | tail -1 |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
    done
  done
fi
