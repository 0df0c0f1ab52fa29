#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
N=6; SW=480; SH=270
head -c $((N*SW*SH*3*2)) /dev/urandom > /tmp/frames_bgr.bin
head -c $((N*SW*SH*3/2*2)) /dev/urandom > /tmp/frames_nv12.bin
for tool in racecheck memcheck synccheck; do
  timeout 45 compute-sanitizer --tool $tool --print-limit 15 examples/stitch_demo $N $SW $SH 1536 4 2 /tmp/frames_bgr.bin /tmp/out_$tool.bin > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -c 'hazard\|Invalid\|Error' gpurun_out/r02_sanitizer_$tool.log) : $(tail -1 gpurun_out/r02_sanitizer_$tool.log)"
done
timeout 40 compute-sanitizer --tool memcheck --print-limit 15 examples/stitch_demo $N $SW $SH 1536 4 2 /tmp/frames_nv12.bin /tmp/out_wire.bin 1280 640 > gpurun_out/r02_sanitizer_memcheck_wire.log 2>&1
echo "memcheck wire rc=$? : $(tail -1 gpurun_out/r02_sanitizer_memcheck_wire.log)"
cmp /tmp/out_racecheck.bin /tmp/out_memcheck.bin && echo "outputs identical under racecheck / memcheck"
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
