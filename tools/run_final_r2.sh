# round-end evidence on one B200: tests, smoke, default bench, reference arm, the other workloads, per-query launch list
bash tools/run_round_checks.sh
bash tools/run_workloads.sh
bash tools/run_localize_launches.sh
