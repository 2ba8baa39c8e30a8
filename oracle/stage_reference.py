#!/usr/bin/env python3
"""Stage a mechanically patched copy of the reference for host (g++) compilation.

TEST INFRASTRUCTURE. Nothing on the product path may depend on what this produces
except the Bifrost *core* scene-handle headers that the drop-in host shim
(`bifrost3d_b200/host/Renderer.cpp`) is compiled against, exactly as it would be inside
the reference's own source tree.

Input : /root/reference (read-only)   [override with $BIFROST_REFERENCE]
Output: baseline/_ref/                (git-ignored; travels to the GPU box with the snapshot)
          Bifrost/            <- core/Bifrost/Bifrost
          OptiXRenderer/      <- extensions/OptiXRenderer/OptiXRenderer
          OptiXRendererTests/ <- tests/OptiXRendererTests
          gtest/              <- extensions/gtest

The reference is MSVC-dialect C++17. The patches below are purely syntactic (they do not
change any arithmetic) and are the recipe recorded in SURVEY.md section 8(c):
  S2  `unsigned int(x)` functional casts  -> `(unsigned int)(x)`
  S3  five one-line dialect fixes (typename, std::min, Math::RGBA, Half.h case)
  S4  test harness: lambda param Material& -> Material, forward declaration
No reference source is committed to this repository.
"""
import os
import re
import shutil
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("BIFROST_REFERENCE", "/root/reference"))
OUT = REPO / "baseline" / "_ref"

COPIES = [
    ("core/Bifrost/Bifrost", "Bifrost"),
    ("extensions/OptiXRenderer/OptiXRenderer", "OptiXRenderer"),
    ("tests/OptiXRendererTests", "OptiXRendererTests"),
    ("extensions/gtest", "gtest"),
    ("extensions/ImageOperations/ImageOperations", "ImageOperations"),
]

CAST_RE = re.compile(r"(?<!operator )\bunsigned (short|int|char)\(")


def patch_text(path: Path, subs):
    text = path.read_text(encoding="latin-1")
    new = text
    for old, rep in subs:
        if old not in new:
            print(f"  [warn] pattern not found in {path}: {old!r}")
        new = new.replace(old, rep)
    if new != text:
        path.write_text(new, encoding="latin-1")


def main():
    if not REF.exists():
        print(f"reference not found at {REF}; nothing staged", file=sys.stderr)
        return 1
    if OUT.exists():
        shutil.rmtree(OUT)
    OUT.mkdir(parents=True)
    for src, dst in COPIES:
        shutil.copytree(REF / src, OUT / dst)

    # S2: MSVC functional casts with multi-word type names.
    n_cast = 0
    for sub in ("Bifrost", "OptiXRenderer", "OptiXRendererTests", "ImageOperations"):
        for p in (OUT / sub).rglob("*"):
            if p.suffix not in (".h", ".cpp", ".impl", ".cu"):
                continue
            text = p.read_text(encoding="latin-1")
            new, n = CAST_RE.subn(r"(unsigned \1)(", text)
            if n:
                p.write_text(new, encoding="latin-1")
                n_cast += n

    # S3: dialect fixes.
    B = OUT / "Bifrost"
    patch_text(B / "Math/Matrix.h", [
        ("    Matrix<R, C, T>::RowType res;\n    for (int c = 0; c < C; ++c)\n        res[c] = dot(lhs, rhs.get_column(c));",
         "    typename Matrix<R, C, T>::RowType res;\n    for (int c = 0; c < C; ++c)\n        res[c] = dot(lhs, rhs.get_column(c));")])
    patch_text(B / "Assets/Mesh.h", [
        ("new std::iterator_traits<RandomAccessIterator>::value_type[",
         "new typename std::iterator_traits<RandomAccessIterator>::value_type[")])
    patch_text(B / "Core/ChangeSet.h", [
        ("int copyable_elements = min(new_size, m_size);", "int copyable_elements = std::min(new_size, m_size);")])
    patch_text(B / "Assets/Image.h", [
        ("            RGBA pixel = get_pixel(image_ID, i);", "            Math::RGBA pixel = get_pixel(image_ID, i);")])
    (B / "Math/Half.h").write_text('#include <Bifrost/Math/half.h>\n')

    # S4: test harness only.
    T = OUT / "OptiXRendererTests"
    for f in ("ShadingModels/DefaultShadingTest.h", "ShadingModels/TransmissiveShadingTest.h"):
        p = T / f
        text = p.read_text(encoding="latin-1")
        new = re.sub(r"\[(.*?)\]\(Material& ", r"[\1](Material ", text)
        p.write_text(new, encoding="latin-1")
    p = T / "BSDFTestUtils.h"
    text = p.read_text(encoding="latin-1")
    marker = "namespace OptiXRenderer::BSDFTestUtils {"
    if marker in text and "float3 w_from_cos_theta(float cos_theta);" not in text:
        text = text.replace(marker, marker + "\ninline optix::float3 w_from_cos_theta(float cos_theta);", 1)
        p.write_text(text, encoding="latin-1")
    else:
        print("  [warn] BSDFTestUtils.h forward declaration not inserted")

    print(f"staged reference into {OUT} ({n_cast} cast rewrites)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
