#!/bin/bash
# static opcode mix of one kernel of liblbgpu.so:  tools/sass_mix.sh <mangled-name-substring> [lib]
lib=${2:-hybird_b200/liblbgpu.so}
fn=$(cuobjdump -sass $lib | grep "Function :" | grep "$1" | head -1 | awk '{print $3}')
echo "kernel: $fn"
cuobjdump -sass -fun "$fn" $lib > /tmp/mix.sass
python - <<'PY'
import re,collections
c=collections.Counter()
for l in open('/tmp/mix.sass'):
    m=re.search(r'/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)',l)
    if m: c[m.group(2).split('.')[0]]+=1
print('static instructions', sum(c.values())); print(c.most_common(40))
PY
