"""Extracts the two border copy tables from the reference's own shader source into tests/golden/border_offsets.json.

Source: /root/reference/Code/Maple/src/Shaders/DDGI/BorderUpdate.glsl:25-133 (`const ivec4 Offsets[68]` under DEPTH_PROBE,
`Offsets[36]` otherwise).  Run in the build container (the reference checkout is not present on the GPU box):
    python tests/golden/make_border_offsets.py
"""
import json
import os
import re

SRC = "/root/reference/Code/Maple/src/Shaders/DDGI/BorderUpdate.glsl"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "border_offsets.json")

text = open(SRC).read()
depth_part, irr_part = text.split("#else", 2)[0], text.split("#else", 2)[1]
# the first '#else' belongs to the PROBE_SIZE block; split on the table declarations instead
m68 = re.search(r"Offsets\[68\]\s*=\s*ivec4\[\]\((.*?)\);", text, re.S)
m36 = re.search(r"Offsets\[36\]\s*=\s*ivec4\[\]\((.*?)\);", text, re.S)
parse = lambda body: [[int(v) for v in t] for t in re.findall(r"ivec4\(\s*(-?\d+)\s*,\s*(-?\d+)\s*,\s*(-?\d+)\s*,\s*(-?\d+)\s*\)", body)]
tables = {"depth": parse(m68.group(1)), "irradiance": parse(m36.group(1)), "source": "Code/Maple/src/Shaders/DDGI/BorderUpdate.glsl:25-133"}
assert len(tables["depth"]) == 68 and len(tables["irradiance"]) == 36
json.dump(tables, open(OUT, "w"))
print("wrote", OUT)
