# launch list of one C1 query (ncu, serialised) next to the un-profiled wall times
timeout 300 python tools/probe_localize.py 20 > gpurun_out/probe_localize_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/localize_launches.csv python tools/probe_localize.py 1 > gpurun_out/probe_localize_ncu.log 2>&1
echo "exit $?"; cat gpurun_out/probe_localize_plain.log | tail -3
