tools/bin/push_bench 256 64 204800
tools/bin/push_bench 256 64 1024000
tools/bin/push_bench 16 256 204800
tools/bin/push_bench 1 1024 204800
