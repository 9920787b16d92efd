"""Cut the device code (kernels, device functions, plain structs / constants) out of a .cu file so that it can be compiled
as ordinary C++ against tests/cpu_shim/cuda_shim.h.  TEST INFRASTRUCTURE ONLY.

Host functions (anything with a <<<...>>> launch, the extern "C" entry points, functions returning an error code) are
dropped; #include lines are dropped (the shim provides what the device code needs); inside cooperative kernels
`__shared__ T name[N];` becomes a per-block array (all blocks of a cooperative launch are resident at once on the CPU too).
"""
import re
import sys

DROP_PREFIX = ("static int ", "extern \"C\"", "#include", "static size_t knn", "static int\n")


def chunks(text):
    """top-level chunks of a translation unit: namespace lines are emitted on their own, every other entity (up to the `;` or
    the closing brace at depth 0 relative to the namespaces) is one chunk"""
    out, cur, depth, ns_depth = [], [], 0, 0
    in_macro = False
    for line in text.splitlines(keepends=True):
        stripped = line.strip()
        if in_macro:                                         # continuation lines of a multi-line #define
            in_macro = stripped.endswith("\\")
            continue
        if depth == ns_depth and not cur:
            if re.match(r"namespace\s+\w*\s*(=|\{)", stripped) and stripped.endswith("{"):
                out.append(("ns", line)); depth += 1; ns_depth += 1
                continue
            if stripped.startswith("}") and "namespace" in stripped and ns_depth > 0:
                out.append(("ns", line)); depth -= 1; ns_depth -= 1
                continue
            if not stripped or stripped.startswith("//"):
                out.append(("c", line))
                continue
            if stripped.startswith("#"):                     # preprocessor lines: dropped (the shim supplies the includes)
                in_macro = stripped.endswith("\\")
                continue
        cur.append(line)
        code = re.sub(r"//.*", "", line)
        code = re.sub(r'"(\\.|[^"\\])*"', '""', code)
        depth += code.count("{") - code.count("}")
        if depth == ns_depth and (code.rstrip().endswith(";") or code.rstrip().endswith("}")):
            out.append(("e", "".join(cur)))
            cur = []
    if cur:
        out.append(("e", "".join(cur)))
    return out


# anything that touches tensor cores, TMA, mbarriers, inline PTX or the CUDA runtime cannot be emulated and is dropped
NOT_EMULATED = re.compile(r"asm\s*(volatile)?\s*\(|CUtensorMap|\btmem_|\bumma_|\btma_|\bmbar_|\btc_fence|\btc_commit|cudaLaunch|cudaStream_t|"
                          r"cudaError_t|__cvta|cudaEvent|cudaFunc|cudaMalloc|cudaMem|cluster_sync|cluster_ctarank|mapa_shared")


# host mode (encoder_emul.cpp): host functions are kept and run against a fake CUDA runtime; only what needs the real
# tensor-core / TMA / mbarrier hardware is dropped, and kernel launches are rewritten into shim launches
NOT_EMULATED_HOST = re.compile(r"asm\s*(volatile)?\s*\(|\btmem_|\bumma_|\btma_load|\btma_prefetch|\bmbar_|\btc_fence|\btc_commit|__cvta|"
                               r"cluster_sync|cluster_ctarank|mapa_shared|cudaLaunchConfig_t|cudaGetDriverEntryPoint")
HOST_MODE = False
# tc mode (attention_emul.cpp): tests/cpu_shim/tc_emul.h models mbarrier / TMA / tcgen05 / TMEM functionally, so kernels that
# use the PTX wrappers are kept; the wrappers themselves (inline asm) and anything cluster-related are dropped
NOT_EMULATED_TC = re.compile(r"asm\s*(volatile)?\s*\(|__cvta|cudaLaunch|cudaStream_t|cudaError_t|cudaEvent|cudaFunc|cudaMalloc|cudaMem")
TC_MODE = False


def rewrite_for_tc(entity):
    entity = re.sub(r'asm volatile\("st\.shared\.v4\.b32[^"]*"\s*::\s*"r"\((.*?)\),\s*"r"\((.*?)\),\s*"r"\((.*?)\),\s*"r"\((.*?)\),\s*"r"\((.*?)\)\s*:\s*"memory"\);',
                    r"shim_st_shared_v4(\1, \2, \3, \4, \5);", entity, flags=re.S)
    return entity.replace("extern __shared__ uint8_t smem_raw[];", "uint8_t *smem_raw = shim::dyn_smem();")


def rewrite_launches(entity):
    """kernel<<<grid, block[, smem[, stream]]>>>(args)  ->  shim_launch(SHIM_CFG(grid, block, ...), kernel, args)"""
    return re.sub(r"([A-Za-z_][\w:]*)\s*<<<(.*?)>>>\s*\(", lambda m: f"shim_launch(SHIM_CFG({m.group(2)}), {m.group(1)}, ", entity, flags=re.S)


def keep(entity):
    head = entity.lstrip()
    if TC_MODE:
        if "<<<" in entity or NOT_EMULATED_TC.search(entity):
            return False
        return not re.match(r"(template\s*<[^>]*>\s*)?static\s+int\b", head) and not head.startswith("extern \"C\"")
    if HOST_MODE:
        return not NOT_EMULATED_HOST.search(entity) and not head.startswith("#include")
    if "<<<" in entity or NOT_EMULATED.search(entity):
        return False
    if re.match(r"(template\s*<[^>]*>\s*)?static\s+int\b", head):
        return False
    if any(head.startswith(p) for p in DROP_PREFIX):
        return False
    if re.match(r"static int\b", head):
        return False
    return True


def per_block_shared(entity):
    """cooperative kernels: `__shared__ float a[N];` -> one array per resident block"""
    if "this_grid()" not in entity:
        return entity
    def repl(m):
        ty, name, dim = m.group(1), m.group(2), m.group(3)
        return f"static {ty} {name}_all[SHIM_MAX_BLOCKS][{dim}]; {ty} *{name} = {name}_all[blockIdx.x];"
    return re.sub(r"__shared__\s+(\w+)\s+(\w+)\[([^\]]+)\];", repl, entity)


def extract(path, host=False, tc=False):
    global HOST_MODE, TC_MODE
    HOST_MODE, TC_MODE = host, tc
    text = open(path).read()
    out = []
    for kind, body in chunks(text):
        if kind in ("ns", "c"):
            out.append(body)
            continue
        if tc:
            body = rewrite_for_tc(body)
        if keep(body):
            body = per_block_shared(body)
            out.append(rewrite_launches(body) if host else body)
    HOST_MODE = TC_MODE = False
    return "".join(out)


if __name__ == "__main__":
    sys.stdout.write(extract(sys.argv[1], host="--host" in sys.argv, tc="--tc" in sys.argv))
