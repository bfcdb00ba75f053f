set -x
mkdir -p gpurun_out
for i in 1 2; do
for p in 1 0; do
LCX_PDL=$p timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-target --algorithm gram --e2e-fit budget > gpurun_out/r02_run35_pdl${p}_$i.json 2> gpurun_out/r02_run35_pdl${p}_$i.err; echo "bench rc=$?"
done; done
