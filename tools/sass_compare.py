"""Per-kernel SASS comparison of two builds of libapg_b200.so (cuobjdump -sass, instruction text hashed per function):
   python tools/sass_compare.py OLD.so NEW.so
Used to show that the GPU-verified kernels are unchanged when new kernels are added to the library."""
import subprocess, re, hashlib, sys
def funcs(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    res, name, buf = {}, None, []
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            if name: res[name] = hashlib.md5("\n".join(buf).encode()).hexdigest()
            name, buf = m.group(1), []
        elif name and "/*" in ln:
            buf.append(re.sub(r"/\*[0-9a-f]{4}\*/", "", ln).strip())
    if name: res[name] = hashlib.md5("\n".join(buf).encode()).hexdigest()
    return res
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
same = [k for k in a if k in b and a[k] == b[k]]
diff = [k for k in a if k in b and a[k] != b[k]]
missing = [k for k in a if k not in b]
print("old kernels:", len(a), "identical in new lib:", len(same), "different:", len(diff), "missing:", len(missing))
for k in diff + missing: print("  !!", k)
print("new-only kernels:", len([k for k in b if k not in a]))
