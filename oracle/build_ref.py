"""oracle/_ref: the reference's own source, compiled where it lies under /root/reference.  TEST INFRASTRUCTURE.

    python -m oracle.build_ref            # -> oracle/_ref/libsplat_ref.so (git-ignored; travels to the GPU box)

The reference's hot path cannot be built the normal way here (OpenGL 4.6 / GLEW / SDL2 / glm / Eigen are absent and
its build is cmake + vcpkg), so this recipe compiles exactly the files that carry the path's arithmetic:
  * host C++   src/gaussiancloud.cpp, src/camerasconfig.cpp, src/ply.cpp, src/core/binaryattribute.cpp, src/core/log.cpp  --
               unmodified, against oracle/ref_shim/{glm/glm.hpp, Eigen/Dense} stand-ins and the REAL nlohmann/json 3.11.3 the image
               carries (cudnn_frontend/thirdparty; oracle/ref_shim/nlohmann/json.hpp is the fallback stand-in);
  * GLSL       shader/presort_compute.glsl, splat_vert.glsl (4 variants: FULL_SH x FRAMEBUFFER_SRGB), splat_geom.glsl,
               splat_frag.glsl -- their text is wrapped into C++ namespaces after the mechanical rewrites below (GLSL
               declarations C++ has no syntax for) and compiled against oracle/ref_shim/glsl_shim.h.
No reference source is copied into the repository: the rewritten shader text is a temporary file that is deleted
after compilation; only the shared library stays in oracle/_ref/.
"""
from __future__ import annotations

import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SPLAT_REFERENCE_DIR", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libsplat_ref.so")

PRELUDE = {
    "compute": "static uvec3 gl_GlobalInvocationID;\n",
    "vertex": "static vec4 gl_Position;\n",
    "geometry": ("struct GlIn { vec4 gl_Position; };\nstatic GlIn gl_in[1];\nstatic vec4 gl_Position;\n"
                 "struct Emitted { vec4 pos, color, cov2inv; vec2 p; };\nstatic Emitted emitted[4];\nstatic int numEmitted;\n"
                 "static void EmitVertex();\nstatic void EndPrimitive() {}\n"),
    "fragment": "static vec4 gl_FragCoord;\nstatic bool gl_discarded;\n",
}
POSTLUDE = {
    "geometry": ("static void EmitVertex()\n{\n    if (numEmitted < 4) { emitted[numEmitted].pos = gl_Position; emitted[numEmitted].color = frag_color;\n"
                 "        emitted[numEmitted].cov2inv = frag_cov2inv; emitted[numEmitted].p = frag_p; }\n    numEmitted++;\n}\n"),
}


def glsl_to_cpp(text: str, stage: str, namespace: str, defines: list[str]) -> str:
    """GLSL declarations that have no C++ spelling are rewritten; every statement and expression is left as is."""
    t = text
    # buffer blocks:  layout(std430, binding = N) readonly buffer Name { T member[]; };  ->  static Buffer<T> member;
    t = re.sub(r"layout\s*\([^)]*\)\s*(?:readonly|writeonly)?\s*buffer\s+\w+\s*\{\s*(\w+)\s+(\w+)\s*\[\s*\]\s*;\s*\}\s*;",
               r"static Buffer<\1> \2;", t)
    # stage layout qualifiers:  layout(points) in;  layout(triangle_strip, max_vertices = 4) out;  layout(local_size_x = 256) in;
    t = re.sub(r"layout\s*\([^)]*\)\s*(?:in|out)\s*;", "", t)
    # layout(binding = 4, offset = 0) uniform atomic_uint output_count;
    t = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+", "static ", t)
    # interface variables
    t = re.sub(r"^[ \t]*in\s+(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"static \1 \2[1];", t, flags=re.M)  # geometry inputs: one point
    t = re.sub(r"^[ \t]*(?:in|out|uniform)\s+(?=\w+\s+\w+\s*;)", "static ", t, flags=re.M)
    # entry point and discard
    t = re.sub(r"\bvoid\s+main\s*\(\s*(?:void)?\s*\)", "void shader_main()", t)
    t = re.sub(r"\bdiscard\s*;", "{ gl_discarded = true; return; }", t)
    head = "".join(f"#define {d}\n" for d in defines)
    tail = "".join(f"#undef {d}\n" for d in defines)
    return (f"namespace {namespace}\n{{\nusing namespace glsl;\n{PRELUDE[stage]}{head}{t}\n{POSTLUDE.get(stage, '')}{tail}}}  // namespace {namespace}\n\n")


VARIANTS = [
    ("presort_compute.glsl", "compute", "ref_presort", []),
    ("splat_vert.glsl", "vertex", "ref_vs_base", []),
    ("splat_vert.glsl", "vertex", "ref_vs_full", ["FULL_SH"]),
    ("splat_vert.glsl", "vertex", "ref_vs_base_srgb", ["FRAMEBUFFER_SRGB"]),
    ("splat_vert.glsl", "vertex", "ref_vs_full_srgb", ["FULL_SH", "FRAMEBUFFER_SRGB"]),
    ("splat_geom.glsl", "geometry", "ref_gs", []),
    ("splat_frag.glsl", "fragment", "ref_fs", []),
]


def real_nlohmann_include() -> str | None:
    """the image carries nlohmann/json 3.11.3 as a third-party header of cudnn_frontend: when it is there the reference's
    camerasconfig.cpp is compiled against the REAL library (the reference pins `nlohmann-json` through vcpkg); the stand-in in
    oracle/ref_shim/nlohmann is only the fallback"""
    import sysconfig

    roots = {sysconfig.get_paths().get("purelib"), sysconfig.get_paths().get("platlib")}
    for r in roots:
        if r:
            d = os.path.join(r, "include", "cudnn_frontend", "thirdparty")
            if os.path.exists(os.path.join(d, "nlohmann", "json.hpp")):
                return d
    return None


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "shader")) and os.path.isdir(os.path.join(REF, "src"))


def build(force: bool = False) -> str | None:
    """Returns the library path, or None when /root/reference is not present (e.g. on the GPU box: the prebuilt
    oracle/_ref/libsplat_ref.so that travelled with the snapshot is used as is)."""
    if not available():
        return LIB if os.path.exists(LIB) else None
    srcs = [os.path.join(REF, "src", f) for f in ("gaussiancloud.cpp", "camerasconfig.cpp", "ply.cpp", os.path.join("core", "binaryattribute.cpp"), os.path.join("core", "log.cpp"))]
    deps = srcs + [os.path.join(REF, "shader", v[0]) for v in VARIANTS] + [os.path.join(HERE, "ref_harness.cpp"), os.path.abspath(__file__),
                                                                            os.path.join(HERE, "ref_shim", "glsl_shim.h"), os.path.join(HERE, "ref_shim", "glm", "glm.hpp"),
                                                                            os.path.join(HERE, "ref_shim", "Eigen", "Dense"), os.path.join(HERE, "ref_shim", "nlohmann", "json.hpp")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="splat_ref_") as tmp:
        gen = os.path.join(tmp, "gen_shaders.inc")
        with open(gen, "w") as f:
            f.write("// generated by oracle/build_ref.py from /root/reference/shader/*.glsl -- temporary, never committed\n")
            for fname, stage, ns, defines in VARIANTS:
                f.write(glsl_to_cpp(open(os.path.join(REF, "shader", fname)).read(), stage, ns, defines))
        real_json = None if os.environ.get("SPLAT_REF_JSON_STANDIN") else real_nlohmann_include()
        json_flags = ["-I", real_json, "-DREF_REAL_NLOHMANN=1"] if real_json else []  # searched before ref_shim: <nlohmann/json.hpp> is the real one
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-march=x86-64-v2",
               "-DNDEBUG", "-w"] + json_flags + ["-I", os.path.join(HERE, "ref_shim"), "-I", os.path.join(REF, "src"), "-I", tmp,
               os.path.join(HERE, "ref_harness.cpp")] + srcs + ["-o", LIB]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("oracle/_ref build failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
