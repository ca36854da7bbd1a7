#!/bin/bash
# runs every diagnostic tool once with its defaults (or small arguments) and reports which ones still work
for t in "config1_s1.py" "config3_check.py" "conv_bench.py 1024 f16" "conv_bench_tails.py" "e2e_probe.py" "metrics_bench.py" "ncu_conv_layer.py 128 0 256 8 3 1 1024" "parity_margin.py" "s1_margin.py 16" "profile_models.py" "probe_conv_accuracy.py" "wasserstein_bench.py" "latency.py" "probe_umma_view.py" "probe_tma_permuted.py"; do
  timeout 240 python tools/$t > /tmp/tool.log 2>&1; rc=$?
  echo "== $t rc=$rc: $(tail -1 /tmp/tool.log | cut -c1-160)"
done
