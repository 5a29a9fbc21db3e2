python tools/latency_probe.py
WFM_NO_FAST_CREATE=1 python tools/latency_probe.py
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
