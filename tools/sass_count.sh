#!/bin/bash
# static SASS statistics of the fp32 restraint kernel: total instructions and opcode histogram
cuobjdump -sass "${1:-trrosettax2-dynamics_b200/build/restraints.o}" 2>/dev/null | awk '/Function : _ZN3trx17restraints_kernelIfE/{f=1} f' | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T] +//' | awk '{split($1,a,"."); c[a[1]]++; n++} END{print "total",n; for(k in c) print c[k],k}' | sort -rn | head -${2:-14}
