cd tools/scratch
./tma_test2 2 64 8 0 0
./tma_test2 2 64 8 16 0
./tma_test2 2 64 8 15 0
./tma_test2 2 48 44 16 0
./tma_test2 2 48 44 15 2
./tma_test2 3 64 8 0 0
./tma_test2 3 48 44 15 2
./tma_test2 3 48 44 16 2
./tma_test2 3 128 32 15 2
