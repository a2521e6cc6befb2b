# usage: r4_report.sh TAG launch [n]   -- top source lines of one k_group_analyse launch of gpurun_out/TAG_analyse.ncu-rep
TAG=$1; L=$2; N=${3:-45}
mkdir -p /tmp/sass_$TAG && cd /tmp/sass_$TAG
[ -f fused.sass ] || { cuobjdump -xelf all /root/repo/wfsim_b200/csrc/libwfsim_b200.so > /dev/null 2>&1; nvdisasm -g -c fused.sm_100a.cubin > fused.sass; }
cd /root/repo
[ -f gpurun_out/${TAG}_source.csv ] || ncu -i gpurun_out/${TAG}_analyse.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
python profiles/sass_top_lines.py gpurun_out/${TAG}_source.csv /tmp/sass_$TAG/fused.sass k_group_analyse $L wfsim_b200/csrc/fused.cu $N
