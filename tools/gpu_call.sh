set -x
mkdir -p gpurun_out
python tools/tfam_ncu_probe.py > gpurun_out/r02_tfam_probe_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tfam_fused -c 4 -o gpurun_out/r02_tfam_fused python tools/tfam_ncu_probe.py > gpurun_out/r02_tfam_probe_ncu.log 2>&1
ls -la gpurun_out/
