#!/bin/bash
mkdir -p gpurun_out
timeout 900 python examples/newnet_dastdp.py --seconds 3600 --out gpurun_out/r2_newnet_dastdp_1h_shist.csv > gpurun_out/r2_newnet_dastdp_1h.txt 2>&1; tail -8 gpurun_out/r2_newnet_dastdp_1h.txt
python - <<'PY'
import numpy as np
a=np.loadtxt('gpurun_out/r2_newnet_dastdp_1h_shist.csv',delimiter=',',skiprows=1)
np.savetxt('gpurun_out/r2_newnet_dastdp_1h_shist_every10s.csv', a[99::100], delimiter=',', header='t_s,s_n1_n2,sd_n1_n2,mean_s', comments='', fmt='%.6g')
print(a.shape, a[-1])
PY
