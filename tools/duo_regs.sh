#!/bin/bash
# development aid: build a model with given consumer/producer register budgets and report spills per region
# usage: tools/duo_regs.sh model CREGS PREGS
cd /root/repo
export JSDE_EXTRA_NVCC_FLAGS="-DJSDE_DUO_CONSUMER_REGS=$2 -DJSDE_DUO_PRODUCER_REGS=$3 $4"
LIB=$(python -c "
from jitcsde_b200 import _build, models
print(_build.build(models.ALL['$1']))")
cuobjdump -sass $LIB | awk '/Function : .*k_integrate.*Lb0/{f=1} /Function : /{if(!/k_integrate.*Lb0/)f=0} f' > /tmp/duo.sass
python - <<'PY'
import re
lines = open('/tmp/duo.sass').read().split('\n')
region = 'pre'; cnt = {}
for l in lines:
    if 'USETMAXREG.DEALLOC' in l: region = 'producer'
    if 'USETMAXREG.TRY_ALLOC' in l: region = 'consumer'
    m = re.search(r'\s(STL|LDL)[\.\s]', l)
    c = cnt.setdefault(region, {'STL':0,'LDL':0,'n':0})
    if re.search(r'/\*[0-9a-f]{4}\*/', l): c['n'] += 1
    if m: c[m.group(1)] += 1
print(cnt)
PY
