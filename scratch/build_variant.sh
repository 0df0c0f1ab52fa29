# usage: bash scratch/build_variant.sh <name> <extra nvcc flags...>   -> scratch/variants/libvsb200_<name>.so
set -e
NAME=$1; shift
D=/tmp/vsb_variant_$NAME; rm -rf $D; mkdir -p $D scratch/variants
cd video-stitcher_b200/csrc
for f in vsb_common vsb_primitives vsb_pipeline vsb_calib; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-w --fmad=false -Xptxas -v "$@" -c $f.cu -o $D/$f.o 2> $D/$f.log &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/variants/libvsb200_$NAME.so $D/*.o -cudart static -ldl
grep -A2 "k_remap_stage[12]_tab\|k_blendENS\|k_down2\|k_coarse" $D/vsb_pipeline.log | grep -E "Compiling|registers|spill" | sed 's/ptxas info    : //'
