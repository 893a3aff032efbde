// ek_processes_m (reference src/processes.f90): terminate, check_master, the process "grid", plus the MPI-free
// rank-per-GPU launcher of this twin.  `mpirun -np P` does not exist in this image; `--ngpu P` forks P-1 ranks
// BEFORE the CUDA runtime is touched (fork after CUDA initialisation is undefined), rank r drives GPU r and the
// 128-byte NCCL id travels through an anonymous shared mapping instead of mpi_bcast.
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <ctype.h>

#include "ek_app.hpp"
#include "launcher.hpp"

namespace ekapp {

int g_block_size = 64;
const char* const g_version = "20160808";
double g_wtime_init = 0.0;

static int s_rank = 0, s_size = 1;
static SharedBoard* s_board = nullptr;

double wtime() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void set_world(int rank, int size) {
  s_rank = rank;
  s_size = size;
}
int world_rank() { return s_rank; }
int world_size() { return s_size; }
bool check_master() { return s_rank == 0; }

// processes.f90:122-139: print the events (master), `[Info]`/`[Error]` + message on stderr, then mpi_abort.
// Here it unwinds to main(), which ends the whole job with that code (the other ranks are killed with it).
[[noreturn]] void terminate(const std::string& error_message, int error_code) { throw Terminate{error_message, error_code}; }

// setup_distribution (processes.f90:17-36) + layout_procs (:56-65) for the B200 backend: the eigenvector matrix
// is dealt by contiguous column slabs, one per GPU, i.e. a 1 x P grid in row-major rank order.
void setup_distribution(ek_process_t& proc) {
  proc.my_rank = s_rank;
  proc.n_procs = s_size;
  proc.n_procs_row = 1;
  proc.n_procs_col = s_size;
  proc.my_proc_row = 0;
  proc.my_proc_col = s_rank;
  proc.context = 0;
  if (proc.my_rank == 0) {  // processes.f90:27-30
    printf("BLACS process grid: %d x %d (%d)\n", proc.n_procs_row, proc.n_procs_col, proc.n_procs);
    fflush(stdout);
  }
}

// ---------------------------------------------------------------- launcher
SharedBoard* shared_board() { return s_board; }

static pid_t s_children[64];
static int s_num_children = 0;

static void on_sigchld(int) {
  // a rank that died with a non-zero status takes the job down (mpi_abort semantics)
  int status = 0;
  pid_t p;
  while ((p = waitpid(-1, &status, WNOHANG)) > 0) {
    bool bad = (WIFEXITED(status) && WEXITSTATUS(status) != 0) || WIFSIGNALED(status);
    if (bad) {
      for (int i = 0; i < s_num_children; ++i)
        if (s_children[i] != p) kill(s_children[i], SIGKILL);
      _exit(WIFEXITED(status) ? WEXITSTATUS(status) : 128 + WTERMSIG(status));
    }
    for (int i = 0; i < s_num_children; ++i)
      if (s_children[i] == p) s_children[i] = s_children[--s_num_children];
  }
}

void launch_ranks(int nranks) {
  if (nranks <= 1) return;
  if (nranks > 64) terminate("launch_ranks: at most 64 ranks", 1);
  void* m = mmap(nullptr, sizeof(SharedBoard), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (m == MAP_FAILED) terminate("launch_ranks: mmap failed", 1);
  s_board = new (m) SharedBoard();
  s_board->nranks = nranks;
  fflush(stdout);
  fflush(stderr);
  struct sigaction sa;
  memset(&sa, 0, sizeof(sa));
  sa.sa_handler = on_sigchld;
  sa.sa_flags = SA_RESTART | SA_NOCLDSTOP;
  sigaction(SIGCHLD, &sa, nullptr);
  const pid_t rank0 = getpid();
  for (int r = 1; r < nranks; ++r) {
    pid_t p = fork();
    if (p < 0) terminate("launch_ranks: fork failed", 1);
    if (p == 0) {
      prctl(PR_SET_PDEATHSIG, SIGKILL);  // a rank never outlives rank 0
      if (getppid() != rank0) _exit(1);  // ... even if rank 0 died before the line above
      signal(SIGCHLD, SIG_DFL);
      s_num_children = 0;
      set_world(r, nranks);
      return;
    }
    s_children[s_num_children++] = p;
  }
  set_world(0, nranks);
}

// ---------------------------------------------------------------- ranks started by mpirun / srun / torchrun
static std::string s_board_path;

static bool env_int(const char* name, long* out) {
  const char* v = getenv(name);
  if (!v || !*v) return false;
  char* end = nullptr;
  long x = strtol(v, &end, 10);
  if (end == v) return false;
  *out = x;
  return true;
}

bool attach_external_ranks() {
  static const char* const kRank[] = {"OMPI_COMM_WORLD_RANK", "PMI_RANK", "PMIX_RANK", "SLURM_PROCID", "RANK"};
  static const char* const kSize[] = {"OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "PMIX_SIZE", "SLURM_NTASKS", "WORLD_SIZE"};
  long rank = -1, size = -1;
  for (int i = 0; i < 5 && (rank < 0 || size < 0); ++i) {
    long r, z;
    if (env_int(kRank[i], &r) && env_int(kSize[i], &z)) { rank = r; size = z; }
  }
  if (size <= 1 || rank < 0 || rank >= size) return false;
  if (size > 64) terminate("attach_external_ranks: at most 64 ranks", 1);
  // one board per launch: EKB200_RENDEZVOUS names it, else the launcher's job id, else the launcher's pid (all ranks
  // of one node are children of the same mpirun / torchrun agent)
  std::string key;
  if (const char* v = getenv("EKB200_RENDEZVOUS")) key = v;  // a name, or a path (contains '/')
  if (key.empty()) {
    for (const char* name : {"TORCHELASTIC_RUN_ID", "PMIX_NAMESPACE", "OMPI_MCA_ess_base_jobid", "SLURM_STEP_ID", "MASTER_PORT"})
      if (const char* v = getenv(name)) { key = std::string(name) + "_" + v; break; }
    key += "_" + std::to_string((long)getppid());
  }
  if (key.find('/') == std::string::npos) {
    for (auto& ch : key)
      if (!isalnum((unsigned char)ch) && ch != '_' && ch != '-') ch = '_';
    s_board_path = "/dev/shm/ekb200_" + key;
  } else {
    s_board_path = key;
  }
  int fd = open(s_board_path.c_str(), O_RDWR | O_CREAT, 0600);
  if (fd < 0) terminate("attach_external_ranks: cannot open " + s_board_path, 1);
  if (ftruncate(fd, sizeof(SharedBoard)) != 0) terminate("attach_external_ranks: ftruncate failed", 1);
  void* m = mmap(nullptr, sizeof(SharedBoard), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (m == MAP_FAILED) terminate("attach_external_ranks: mmap failed", 1);
  s_board = reinterpret_cast<SharedBoard*>(m);  // a fresh file is all zeros: a valid initial board
  set_world((int)rank, (int)size);
  return true;
}

// Sense-reversing barrier over the shared board (mpi_barrier).
void world_barrier() {
  if (!s_board || s_size <= 1) return;
  const int gen = s_board->barrier_gen.load();
  if (s_board->barrier_count.fetch_add(1) + 1 == s_size) {
    s_board->barrier_count.store(0);
    s_board->barrier_gen.fetch_add(1);
  } else {
    while (s_board->barrier_gen.load() == gen) usleep(50);
  }
}

// Rank 0 waits for the others before it leaves (mpi_finalize).
void finalize_ranks() {
  if (s_size <= 1) return;
  if (!s_board_path.empty()) {  // external launcher: it reaps the ranks; rank 0 removes the board of this launch
    if (s_rank == 0) unlink(s_board_path.c_str());
    return;
  }
  if (s_rank != 0) return;
  sigset_t block;
  sigemptyset(&block);
  sigaddset(&block, SIGCHLD);
  sigprocmask(SIG_BLOCK, &block, nullptr);  // the handler no longer races with this loop
  signal(SIGCHLD, SIG_DFL);
  sigprocmask(SIG_UNBLOCK, &block, nullptr);
  for (int i = 0; i < s_num_children; ++i) {
    int status = 0;
    if (waitpid(s_children[i], &status, 0) > 0) {
      if ((WIFEXITED(status) && WEXITSTATUS(status) != 0) || WIFSIGNALED(status)) {
        fprintf(stderr, "[Error] a rank ended abnormally\n");
        _exit(1);
      }
    }
  }
  s_num_children = 0;
}

void abort_ranks() {
  if (s_rank == 0)
    for (int i = 0; i < s_num_children; ++i) kill(s_children[i], SIGKILL);
  // external launcher: it tears the other ranks down when this one exits non-zero; do not leave the board behind
  if (s_rank == 0 && !s_board_path.empty()) unlink(s_board_path.c_str());
}

}  // namespace ekapp
