"""Regenerate the measured-results tables of README.md and DESIGN.md from the committed records (profiles/r2):
python tools/fill_tables.py     (idempotent: replaces the block between the BEGIN/END markers)
python tools/fill_tables.py --check     exit status 1 if a document's block differs from the records (tests/test_host.py)"""
import re
import subprocess
import sys

table = subprocess.run([sys.executable, "tools/results_table.py", "profiles/r2"], capture_output=True, text=True).stdout.strip()
block = "<!-- BEGIN RESULTS (tools/fill_tables.py) -->\n" + table + "\n<!-- END RESULTS -->"
check = "--check" in sys.argv[1:]
stale = []
for f in ("README.md", "DESIGN.md"):
    s = open(f).read()
    if check:
        if block not in s:
            stale.append(f)
        continue
    if "RESULTS_TABLE_PLACEHOLDER" in s:
        s = s.replace("RESULTS_TABLE_PLACEHOLDER", block)
    else:
        s = re.sub(r"<!-- BEGIN RESULTS \(tools/fill_tables.py\) -->.*?<!-- END RESULTS -->", lambda m: block, s, flags=re.S)
    open(f, "w").write(s)
if check:
    if stale:
        print("stale result tables:", ", ".join(stale))
        sys.exit(1)
    print("result tables match profiles/r2")
else:
    print(table)
