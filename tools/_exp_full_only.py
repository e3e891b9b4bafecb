import sys
sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo")
import scan_experiments as se
se.VARIANTS = {"full": []}
se.run()
