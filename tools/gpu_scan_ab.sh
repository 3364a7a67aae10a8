cd $GRAFT_REPO_ROOT
python tools/scan_c2.py
for f in mamba_tts_project_b200/libmtts_*.so; do MTTS_LIB=$GRAFT_REPO_ROOT/$f python tools/scan_c2.py; done
