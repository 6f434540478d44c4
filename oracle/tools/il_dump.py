#!/usr/bin/env python3
"""Minimal ECMA-335 (.NET) metadata + IL reader — TEST INFRASTRUCTURE ONLY.

Purpose: the reference's variant q-score calls MathNet.Numerics 4.5.1
(`Poisson.CumulativeDistribution`, `Poisson.ProbabilityLn`, `SpecialFunctions.GammaLowerRegularized`,
`GammaLn`, `FactorialLn`, `Binomial.CumulativeDistribution`, `BetaRegularized`), whose *source* is not under
/root/reference. The compiled IL is, inside binaries/5.2.11.163/Pisces_5.2.11.163.tar.gz (MathNet.Numerics.dll).
This tool disassembles named methods from that DLL so the oracle's restatement of the MathNet arithmetic can be
checked instruction-by-instruction against what the reference actually ships (operation order, constants,
loop conditions), instead of relying on memory of the MathNet sources.

Usage:  python oracle/tools/il_dump.py <assembly.dll> <TypeName> <MethodName> [<TypeName> <MethodName> ...]
It never runs on the GPU box and nothing in the product imports it.
"""
import struct, sys

class PE:
    def __init__(self, data):
        self.d = data
        pe = struct.unpack_from('<I', data, 0x3c)[0]
        assert data[pe:pe+4] == b'PE\0\0'
        nsec = struct.unpack_from('<H', data, pe+6)[0]
        optsz = struct.unpack_from('<H', data, pe+20)[0]
        opt = pe + 24
        magic = struct.unpack_from('<H', data, opt)[0]
        ddoff = opt + (96 if magic == 0x10b else 112)
        self.cli_rva, self.cli_sz = struct.unpack_from('<II', data, ddoff + 14*8)
        self.secs = []
        so = opt + optsz
        for i in range(nsec):
            vs, va, rs, ro = struct.unpack_from('<IIII', data, so + i*40 + 8)
            self.secs.append((va, max(vs, rs), ro))
    def off(self, rva):
        for va, sz, ro in self.secs:
            if va <= rva < va + sz:
                return rva - va + ro
        raise ValueError(hex(rva))

# table schemas: list of column kinds
# 's' string idx, 'g' guid idx, 'b' blob idx, 'u2','u4' fixed, ('t',n) simple table idx, ('c',name) coded idx
CODED = {
 'TypeDefOrRef': (2, [0x02, 0x01, 0x1b]),
 'HasConstant': (2, [0x04, 0x08, 0x17]),
 'HasCustomAttribute': (5, [0x06,0x04,0x01,0x02,0x08,0x09,0x0a,0x00,0x0e,0x17,0x14,0x11,0x1a,0x1b,0x20,0x23,0x26,0x27,0x28,0x2a,0x2c,0x2b]),
 'HasFieldMarshal': (1, [0x04, 0x08]),
 'HasDeclSecurity': (2, [0x02, 0x06, 0x20]),
 'MemberRefParent': (3, [0x02, 0x01, 0x1a, 0x06, 0x1b]),
 'HasSemantics': (1, [0x14, 0x17]),
 'MethodDefOrRef': (1, [0x06, 0x0a]),
 'MemberForwarded': (1, [0x04, 0x06]),
 'Implementation': (2, [0x26, 0x23, 0x27]),
 'CustomAttributeType': (3, [None, None, 0x06, 0x0a, None]),
 'ResolutionScope': (2, [0x00, 0x1a, 0x23, 0x01]),
 'TypeOrMethodDef': (1, [0x02, 0x06]),
}
T = lambda n: ('t', n)
C = lambda n: ('c', n)
SCHEMA = {
 0x00: ['u2','s','g','g','g'],
 0x01: [C('ResolutionScope'),'s','s'],
 0x02: ['u4','s','s',C('TypeDefOrRef'),T(0x04),T(0x06)],
 0x03: [T(0x04)],
 0x04: ['u2','s','b'],
 0x05: [T(0x06)],
 0x06: ['u4','u2','u2','s','b',T(0x08)],
 0x07: [T(0x08)],
 0x08: ['u2','u2','s'],
 0x09: [T(0x02),C('TypeDefOrRef')],
 0x0a: [C('MemberRefParent'),'s','b'],
 0x0b: ['u2',C('HasConstant'),'b'],
 0x0c: [C('HasCustomAttribute'),C('CustomAttributeType'),'b'],
 0x0d: [C('HasFieldMarshal'),'b'],
 0x0e: ['u2',C('HasDeclSecurity'),'b'],
 0x0f: ['u2','u4',T(0x02)],
 0x10: ['u4',T(0x04)],
 0x11: ['b'],
 0x12: [T(0x02),T(0x14)],
 0x13: [T(0x14)],
 0x14: ['u2','s',C('TypeDefOrRef')],
 0x15: [T(0x02),T(0x17)],
 0x16: [T(0x17)],
 0x17: ['u2','s','b'],
 0x18: ['u2',T(0x06),C('HasSemantics')],
 0x19: [T(0x02),C('MethodDefOrRef'),C('MethodDefOrRef')],
 0x1a: ['s'],
 0x1b: ['b'],
 0x1c: ['u2',C('MemberForwarded'),'s',T(0x1a)],
 0x1d: ['u4',T(0x04)],
 0x1e: ['u4','u4'],
 0x1f: ['u4'],
 0x20: ['u4','u2','u2','u2','u2','u4','b','s','s'],
 0x21: ['u4'],
 0x22: ['u4','u4','u4'],
 0x23: ['u2','u2','u2','u2','u4','b','s','s','b'],
 0x24: ['u4',T(0x23)],
 0x25: ['u4','u4','u4',T(0x23)],
 0x26: ['u4','s','b'],
 0x27: ['u4','u4','s','s',C('Implementation')],
 0x28: ['u4','u4','s',C('Implementation')],
 0x29: [T(0x02),T(0x02)],
 0x2a: ['u2','u2',C('TypeOrMethodDef'),'s'],
 0x2b: [C('MethodDefOrRef'),'b'],
 0x2c: [T(0x2a),C('TypeDefOrRef')],
}

class Meta:
    def __init__(self, path):
        self.data = open(path, 'rb').read()
        self.pe = PE(self.data)
        cli = self.pe.off(self.pe.cli_rva)
        md_rva, md_sz = struct.unpack_from('<II', self.data, cli + 8)
        self.md = self.pe.off(md_rva)
        d = self.data; p = self.md
        assert struct.unpack_from('<I', d, p)[0] == 0x424A5342
        vlen = struct.unpack_from('<I', d, p+12)[0]
        p += 16 + vlen
        nstreams = struct.unpack_from('<H', d, p+2)[0]
        p += 4
        self.streams = {}
        for _ in range(nstreams):
            o, s = struct.unpack_from('<II', d, p); p += 8
            e = d.index(b'\0', p); name = d[p:e].decode(); p = (e + 4) & ~3
            self.streams[name] = (self.md + o, s)
        self._tables()
    def string(self, i):
        o = self.streams['#Strings'][0] + i
        return self.data[o:self.data.index(b'\0', o)].decode('utf8', 'replace')
    def _tables(self):
        d = self.data
        p = self.streams.get('#~', self.streams.get('#-'))[0]
        heap = d[p+6]
        self.ssz = 4 if heap & 1 else 2; self.gsz = 4 if heap & 2 else 2; self.bsz = 4 if heap & 4 else 2
        valid = struct.unpack_from('<Q', d, p+8)[0]
        p += 24
        self.rows = {}
        for t in range(64):
            if valid >> t & 1:
                self.rows[t] = struct.unpack_from('<I', d, p)[0]; p += 4
        def colsize(c):
            if c == 'u2': return 2
            if c == 'u4': return 4
            if c == 's': return self.ssz
            if c == 'g': return self.gsz
            if c == 'b': return self.bsz
            if c[0] == 't': return 4 if self.rows.get(c[1], 0) >= 65536 else 2
            bits, tabs = CODED[c[1]]
            mx = max(self.rows.get(t, 0) for t in tabs if t is not None)
            return 4 if mx >= (1 << (16 - bits)) else 2
        self.toff = {}; self.rsz = {}; self.csz = {}
        for t in sorted(self.rows):
            cs = [colsize(c) for c in SCHEMA[t]]
            self.csz[t] = cs; self.rsz[t] = sum(cs); self.toff[t] = p
            p += self.rsz[t] * self.rows[t]
    def row(self, t, i):  # 1-based
        p = self.toff[t] + (i-1) * self.rsz[t]
        out = []
        for sz in self.csz[t]:
            out.append(struct.unpack_from('<H' if sz == 2 else '<I', self.data, p)[0]); p += sz
        return out
    def typedefs(self):
        n = self.rows[0x02]
        for i in range(1, n+1):
            r = self.row(0x02, i)
            mstart = r[5]
            mend = self.row(0x02, i+1)[5] if i < n else self.rows[0x06] + 1
            yield i, self.string(r[2]), self.string(r[1]), mstart, mend
    def token_name(self, tok):
        t, i = tok >> 24, tok & 0xffffff
        try:
            if t == 0x06:
                r = self.row(0x06, i)
                owner = ''
                for _, ns, nm, ms, me in self.typedefs():
                    if ms <= i < me: owner = nm
                return f'{owner}::{self.string(r[3])}'
            if t == 0x0a:
                r = self.row(0x0a, i)
                bits, tabs = CODED['MemberRefParent']
                pt, pi = tabs[r[0] & ((1 << bits) - 1)], r[0] >> bits
                pn = '?'
                if pt == 0x01: pn = self.string(self.row(0x01, pi)[2])
                elif pt == 0x02: pn = self.string(self.row(0x02, pi)[1])
                return f'{pn}::{self.string(r[1])}'
            if t == 0x04:
                return 'field ' + self.string(self.row(0x04, i)[1])
            if t == 0x01: return self.string(self.row(0x01, i)[2])
            if t == 0x02: return self.string(self.row(0x02, i)[1])
            if t == 0x2b: return 'methodspec->' + self.token_name(((0x06, 0x0a)[self.row(0x2b, i)[0] & 1] << 24) | (self.row(0x2b, i)[0] >> 1))
        except Exception as e:
            return f'<{e}>'
        return hex(tok)

ONE = {0x00:'nop',0x02:'ldarg.0',0x03:'ldarg.1',0x04:'ldarg.2',0x05:'ldarg.3',0x06:'ldloc.0',0x07:'ldloc.1',0x08:'ldloc.2',0x09:'ldloc.3',
 0x0a:'stloc.0',0x0b:'stloc.1',0x0c:'stloc.2',0x0d:'stloc.3',0x14:'ldnull',0x15:'ldc.i4.m1',0x16:'ldc.i4.0',0x17:'ldc.i4.1',0x18:'ldc.i4.2',
 0x19:'ldc.i4.3',0x1a:'ldc.i4.4',0x1b:'ldc.i4.5',0x1c:'ldc.i4.6',0x1d:'ldc.i4.7',0x1e:'ldc.i4.8',0x25:'dup',0x26:'pop',0x2a:'ret',
 0x58:'add',0x59:'sub',0x5a:'mul',0x5b:'div',0x5c:'div.un',0x5d:'rem',0x5e:'rem.un',0x5f:'and',0x60:'or',0x61:'xor',0x62:'shl',0x63:'shr',0x64:'shr.un',
 0x65:'neg',0x66:'not',0x67:'conv.i1',0x68:'conv.i2',0x69:'conv.i4',0x6a:'conv.i8',0x6b:'conv.r4',0x6c:'conv.r8',0x6d:'conv.u4',0x6e:'conv.u8',
 0x76:'conv.r.un',0x8e:'ldlen',0x90:'ldelem.i1',0x91:'ldelem.u1',0x92:'ldelem.i2',0x93:'ldelem.u2',0x94:'ldelem.i4',0x95:'ldelem.u4',0x96:'ldelem.i8',
 0x97:'ldelem.i',0x98:'ldelem.r4',0x99:'ldelem.r8',0x9a:'ldelem.ref',0x9b:'stelem.i',0x9c:'stelem.i1',0x9d:'stelem.i2',0x9e:'stelem.i4',0x9f:'stelem.i8',
 0xa0:'stelem.r4',0xa1:'stelem.r8',0xa2:'stelem.ref',0xd1:'conv.u2',0xd2:'conv.u1',0xd3:'conv.i',0xe0:'conv.u',0x7a:'throw',0xdc:'endfinally',
 0xb7:'conv.ovf.i4',0xb9:'conv.ovf.i8',0xd6:'add.ovf',0xd8:'mul.ovf',0xda:'sub.ovf'}
BR1 = {0x2b:'br.s',0x2c:'brfalse.s',0x2d:'brtrue.s',0x2e:'beq.s',0x2f:'bge.s',0x30:'bgt.s',0x31:'ble.s',0x32:'blt.s',0x33:'bne.un.s',0x34:'bge.un.s',0x35:'bgt.un.s',0x36:'ble.un.s',0x37:'blt.un.s',0xde:'leave.s'}
BR4 = {0x38:'br',0x39:'brfalse',0x3a:'brtrue',0x3b:'beq',0x3c:'bge',0x3d:'bgt',0x3e:'ble',0x3f:'blt',0x40:'bne.un',0x41:'bge.un',0x42:'bgt.un',0x43:'ble.un',0x44:'blt.un',0xdd:'leave'}
U1 = {0x0e:'ldarg.s',0x0f:'ldarga.s',0x10:'starg.s',0x11:'ldloc.s',0x12:'ldloca.s',0x13:'stloc.s',0x1f:'ldc.i4.s'}
TOK = {0x28:'call',0x6f:'callvirt',0x73:'newobj',0x7b:'ldfld',0x7c:'ldflda',0x7d:'stfld',0x7e:'ldsfld',0x7f:'ldsflda',0x80:'stsfld',0x8d:'newarr',
 0x8c:'box',0xa5:'unbox.any',0x74:'castclass',0x75:'isinst',0xd0:'ldtoken',0x72:'ldstr',0xa3:'ldelem',0xa4:'stelem',0x8f:'ldelema',0x70:'cpobj',0x71:'ldobj',0x81:'stobj',0x27:'calli',0x29:'jmp'}
FE = {0x01:'ceq',0x02:'cgt',0x03:'cgt.un',0x04:'clt',0x05:'clt.un',0x16:'constrained.',0x15:'initobj',0x1e:'readonly.',0x06:'ldftn',0x1a:'rethrow'}

def disasm(m, rva):
    d = m.data; p = m.pe.off(rva)
    h = d[p]
    if h & 3 == 2:
        size = h >> 2; code = p + 1
    else:
        flags, maxstack, size, loc = struct.unpack_from('<HHII', d, p); code = p + 12
    out = []; i = 0
    while i < size:
        o = d[code+i]; a = i; i += 1
        if o in ONE: s = ONE[o]
        elif o in BR1: s = f'{BR1[o]} IL_{i+1+struct.unpack_from("<b", d, code+i)[0]:04x}'; i += 1
        elif o in BR4: s = f'{BR4[o]} IL_{i+4+struct.unpack_from("<i", d, code+i)[0]:04x}'; i += 4
        elif o in U1: s = f'{U1[o]} {struct.unpack_from("<b", d, code+i)[0]}'; i += 1
        elif o == 0x20: s = f'ldc.i4 {struct.unpack_from("<i", d, code+i)[0]}'; i += 4
        elif o == 0x21: s = f'ldc.i8 {struct.unpack_from("<q", d, code+i)[0]}'; i += 8
        elif o == 0x22: s = f'ldc.r4 {struct.unpack_from("<f", d, code+i)[0]!r}'; i += 4
        elif o == 0x23: s = f'ldc.r8 {struct.unpack_from("<d", d, code+i)[0]!r}'; i += 8
        elif o in TOK:
            tok = struct.unpack_from('<I', d, code+i)[0]; i += 4
            s = f'{TOK[o]} {m.token_name(tok)}'
        elif o == 0x45:
            n = struct.unpack_from('<I', d, code+i)[0]; i += 4
            tg = struct.unpack_from(f'<{n}i', d, code+i); i += 4*n
            s = 'switch ' + ','.join(f'IL_{i+t:04x}' for t in tg)
        elif o == 0xfe:
            o2 = d[code+i]; i += 1
            if o2 in (0x16, 0x15, 0x06): 
                tok = struct.unpack_from('<I', d, code+i)[0]; i += 4
                s = f'{FE[o2]} {m.token_name(tok)}'
            elif o2 in (0x09,0x0a,0x0b,0x0c,0x0d,0x0e):
                s = f'{ {0x09:"ldarg",0x0a:"ldarga",0x0b:"starg",0x0c:"ldloc",0x0d:"ldloca",0x0e:"stloc"}[o2]} {struct.unpack_from("<H", d, code+i)[0]}'; i += 2
            else: s = FE.get(o2, f'fe{o2:02x}')
        else: s = f'?? {o:02x}'
        out.append(f'  IL_{a:04x}: {s}')
    return out

def field_doubles(m, name, n):
    """Read n doubles from the FieldRVA blob backing static array initialiser `name`."""
    for i in range(1, m.rows[0x1d] + 1):
        rva, fld = m.row(0x1d, i)
        if m.string(m.row(0x04, fld)[1]) == name:
            return struct.unpack_from(f'<{n}d', m.data, m.pe.off(rva))
    raise KeyError(name)

def main():
    m = Meta(sys.argv[1])
    if sys.argv[2] == '--field-doubles':
        for v in field_doubles(m, sys.argv[3], int(sys.argv[4])): print(repr(v))
        return
    want = list(zip(sys.argv[2::2], sys.argv[3::2]))
    for ti, ns, nm, ms, me in m.typedefs():
        for wt, wm in want:
            if nm != wt: continue
            for mi in range(ms, me):
                r = m.row(0x06, mi)
                if m.string(r[3]) == wm and r[0]:
                    print(f'== {ns}.{nm}::{wm}  (methoddef {mi}, rva 0x{r[0]:x})')
                    print('\n'.join(disasm(m, r[0])))
if __name__ == '__main__':
    main()
