"""Only the writer functions of the reference's `footprint_tools.cli` package (cli/utils.py:86-214) — the text
output right after the scoring path. The click commands themselves are out of scope."""
