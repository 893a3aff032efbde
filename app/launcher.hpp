// MPI-free rank plumbing of ekb200_app (processes.cpp): forked ranks + a shared board for the NCCL id,
// a barrier and small broadcasts (the roles of mpi_bcast / mpi_barrier in the reference's main.f90).
#pragma once
#include <atomic>

namespace ekapp {

struct SharedBoard {
  int nranks = 1;
  std::atomic<int> barrier_count{0};
  std::atomic<int> barrier_gen{0};
  std::atomic<int> id_ready{0};
  unsigned char nccl_id[128] = {0};
  std::atomic<int> error_code{0};
};

SharedBoard* shared_board();
void launch_ranks(int nranks);   // after this call world_rank()/world_size() are set in every process
// Ranks started by an external launcher (mpirun / srun / torchrun --no-python: RANK + WORLD_SIZE, OMPI_COMM_WORLD_*,
// PMI_*, SLURM_PROCID + SLURM_NTASKS in the environment): attaches this process to a file-backed board under /dev/shm
// shared by the ranks of the launch.  Returns false when the environment names no multi-rank launch.
bool attach_external_ranks();
void world_barrier();
void finalize_ranks();
void abort_ranks();

}  // namespace ekapp
