python tools/latency_probe.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2i_lat_launches.csv python tools/latency_probe.py > /dev/null 2>&1
tail -25 gpurun_out/r2i_lat_launches.csv | cut -c1-220
