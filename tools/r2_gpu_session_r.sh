# full GPU suite on the current tree (all pairs now also at PAL and UHD; frame lists), then the window kernel's two forms
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2r_tests.log
python tools/tcv_probe.py --only clip > gpurun_out/r2r_tcv.txt 2>&1
python tools/tcv_probe.py --only "reduce 1x2" >> gpurun_out/r2r_tcv.txt 2>&1
python tools/tcv_probe.py --only gamma >> gpurun_out/r2r_tcv.txt 2>&1
