# one full capture of the draw + P3P kernel of a C1 query's second wave (1 GPU)
CMD="python tools/probe_resect_trace.py"
timeout 300 $CMD > gpurun_out/plain_p3p.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:p3p_draw -s 41 -c 1 -o gpurun_out/r2_p3p $CMD > gpurun_out/ncu_p3p.log 2>&1
echo "full exit $?"; tail -3 gpurun_out/ncu_p3p.log
