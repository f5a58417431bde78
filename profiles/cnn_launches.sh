mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cnn.csv python profiles/cnn_time.py > gpurun_out/cnn_time_under_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/launches_cnn.csv') if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:60], float(r[vi].replace(',',''))/1e3) for r in rows[1:]]
# print the launches of the first mode-0 updateFeatures after setWeights: find first conv1_split
idx=[i for i,(k,_) in enumerate(seq) if 'conv1_split' in k]
i0=idx[2] if len(idx)>2 else idx[0]
for k,v in seq[i0:i0+8]: print(f"{k:62s} {v:8.1f} us")
PY
