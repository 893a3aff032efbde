// MPI-free rank plumbing of ekb200_app (processes.cpp): forked ranks + a shared board for the NCCL id,
// a barrier and small broadcasts (the roles of mpi_bcast / mpi_barrier in the reference's main.f90).
#pragma once
#include <atomic>

namespace ekapp {

struct SharedBoard {
  std::atomic<unsigned> magic{0};   // file-backed boards: set LAST by the creating rank (0); 0 = not initialised yet
  int owner_pid = 0;                // pid of the rank that created the board (a board whose creator is dead is stale)
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
// shared by the ranks of the launch.  OPT-IN: only when EKB200_EXTERNAL_RANKS=1 (or EKB200_RENDEZVOUS names the board) --
// launcher variables alone are also the normal environment of an sbatch script that runs the binary once, without
// srun.  Single node only.  Returns false when not asked for or when the environment names no multi-rank launch.
bool attach_external_ranks();
void world_barrier();
void wait_for_nccl_id();  // ranks > 0: until rank 0 has stored the id on the board
void finalize_ranks();
void abort_ranks();

}  // namespace ekapp
