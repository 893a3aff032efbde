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
void world_barrier();
void finalize_ranks();
void abort_ranks();

}  // namespace ekapp
