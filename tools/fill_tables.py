"""Regenerate the measured-results tables of README.md and DESIGN.md from the committed records (profiles/r2):
python tools/fill_tables.py     (idempotent: replaces the block between the BEGIN/END markers)"""
import re
import subprocess
import sys

table = subprocess.run([sys.executable, "tools/results_table.py", "profiles/r2"], capture_output=True, text=True).stdout.strip()
block = "<!-- BEGIN RESULTS (tools/fill_tables.py) -->\n" + table + "\n<!-- END RESULTS -->"
for f in ("README.md", "DESIGN.md"):
    s = open(f).read()
    if "RESULTS_TABLE_PLACEHOLDER" in s:
        s = s.replace("RESULTS_TABLE_PLACEHOLDER", block)
    else:
        s = re.sub(r"<!-- BEGIN RESULTS \(tools/fill_tables.py\) -->.*?<!-- END RESULTS -->", lambda m: block, s, flags=re.S)
    open(f, "w").write(s)
print(table)
