#!/bin/bash
# rebuild libovdet.so and dump the SASS of one object: tools/sass.sh sim_fused_sm100
cd /root/repo && python -c "from ovdet import build; build.build()" >/dev/null || exit 1
cuobjdump -sass real-time*/csrc/_obj/$1.o | grep -E "^\s+/\*[0-9a-f]{4}\*/|Function" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s+\/\*.*$//' > /tmp/$1.sass
echo "/tmp/$1.sass: $(wc -l < /tmp/$1.sass) lines"
