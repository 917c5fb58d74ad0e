run() { env "$@" timeout 120 python tools/krylov_consistency.py 2>&1 | grep -E "^T|rror" ; }
run TAG=T NN=15728641
run TAG=T NN=15728641 BK2_E=7
run TAG=T NN=15728640 BK2_E=8
run TAG=T NN=15730688 BK2_E=8
run TAG=T NN=3932161 BK2_E=8
run TAG=T NN=1212417 BK2_E=8
run TAG=T NN=6000000 BK2_E=8
run TAG=T NN=6000000 BK2_E=6
