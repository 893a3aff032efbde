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
#include <errno.h>
#include <sys/prctl.h>
#include <sys/stat.h>
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

static constexpr unsigned kBoardMagic = 0x454b4232u;  // "EKB2"

static double rendezvous_timeout() {
  long t = 120;
  env_int("EKB200_RENDEZVOUS_TIMEOUT", &t);
  return t > 0 ? (double)t : 120.0;
}

[[noreturn]] static void rendezvous_failed(const std::string& what) {
  if (s_rank == 0 && !s_board_path.empty()) unlink(s_board_path.c_str());
  terminate("rendezvous of " + std::to_string(s_size) + " ranks timed out (" + what + "): rank " + std::to_string(s_rank) +
                " waited " + std::to_string((long)rendezvous_timeout()) + " s on " + s_board_path +
                " -- were all ranks started on this node (mpirun / srun / torchrun)?  EKB200_RENDEZVOUS_TIMEOUT sets the deadline",
            1);
}

bool attach_external_ranks() {
  // opt-in (an sbatch script running the binary once still has SLURM_PROCID / SLURM_NTASKS in its environment)
  const char* optin = getenv("EKB200_EXTERNAL_RANKS");
  const char* named = getenv("EKB200_RENDEZVOUS");
  if (!((optin && atoi(optin) == 1) || (named && *named))) return false;
  static const char* const kRank[] = {"OMPI_COMM_WORLD_RANK", "PMI_RANK", "PMIX_RANK", "SLURM_PROCID", "RANK"};
  static const char* const kSize[] = {"OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "PMIX_SIZE", "SLURM_NTASKS", "WORLD_SIZE"};
  static const char* const kLocal[] = {"OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS", "", "SLURM_NTASKS_PER_NODE",
                                       "LOCAL_WORLD_SIZE"};
  long rank = -1, size = -1, local = -1;
  for (int i = 0; i < 5 && (rank < 0 || size < 0); ++i) {
    long r, z;
    if (env_int(kRank[i], &r) && env_int(kSize[i], &z)) {
      rank = r;
      size = z;
      long l;
      if (*kLocal[i] && env_int(kLocal[i], &l)) local = l;
    }
  }
  if (size <= 1 || rank < 0 || rank >= size) return false;
  if (size > 64) terminate("attach_external_ranks: at most 64 ranks", 1);
  // the board lives in this node's /dev/shm and the B200 solvers use the GPUs of ONE box
  if (local > 0 && local != size)
    terminate("attach_external_ranks: " + std::to_string(size) + " ranks but only " + std::to_string(local) +
                  " on this node: the B200 solvers run on the GPUs of one node",
              1);
  // one board per launch: EKB200_RENDEZVOUS names it, else the launcher's job id, else the launcher's pid (all ranks
  // of one node are children of the same mpirun / torchrun agent)
  std::string key;
  if (named) key = named;  // a name, or a path (contains '/')
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
  set_world((int)rank, (int)size);
  const double deadline = wtime() + rendezvous_timeout();
  if (rank == 0) {
    // the creating rank: a fresh, exclusively created file (never follows a planted symlink, never reuses the
    // contents a crashed launch left behind), initialised before the magic is published
    int fd = -1;
    for (int attempt = 0; attempt < 8 && fd < 0; ++attempt) {
      fd = open(s_board_path.c_str(), O_RDWR | O_CREAT | O_EXCL | O_NOFOLLOW | O_CLOEXEC, 0600);
      if (fd < 0 && errno == EEXIST) unlink(s_board_path.c_str());  // stale board of an earlier launch
      else if (fd < 0) break;
    }
    if (fd < 0) terminate("attach_external_ranks: cannot create " + s_board_path + ": " + strerror(errno), 1);
    if (ftruncate(fd, sizeof(SharedBoard)) != 0) terminate("attach_external_ranks: ftruncate failed", 1);
    void* m = mmap(nullptr, sizeof(SharedBoard), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) terminate("attach_external_ranks: mmap failed", 1);
    s_board = new (m) SharedBoard();  // barrier fields, id_ready and error_code start from zero
    s_board->nranks = (int)size;
    s_board->owner_pid = (int)getpid();
    s_board->magic.store(kBoardMagic);
    return true;
  }
  // the other ranks: wait for a board that is ours (regular file owned by this user, created by a LIVE rank 0 for
  // the same number of ranks); anything else is a leftover that rank 0 is about to replace -- look again
  for (;;) {
    int fd = open(s_board_path.c_str(), O_RDWR | O_NOFOLLOW | O_CLOEXEC);
    if (fd >= 0) {
      struct stat st;
      bool ok = fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_uid == geteuid() &&
                (size_t)st.st_size >= sizeof(SharedBoard);
      void* m = ok ? mmap(nullptr, sizeof(SharedBoard), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0) : MAP_FAILED;
      close(fd);
      if (m != MAP_FAILED) {
        SharedBoard* b = reinterpret_cast<SharedBoard*>(m);
        if (b->magic.load() == kBoardMagic && b->nranks == (int)size && b->owner_pid > 0 && kill(b->owner_pid, 0) == 0) {
          s_board = b;
          return true;
        }
        munmap(m, sizeof(SharedBoard));
      }
    }
    if (wtime() > deadline) rendezvous_failed("no board from rank 0");
    usleep(2000);
  }
}

// Sense-reversing barrier over the shared board (mpi_barrier).  File-backed boards (external launcher) carry a
// deadline: a rank that never arrives ends the job with a message instead of a silent hang.
void world_barrier() {
  if (!s_board || s_size <= 1) return;
  const bool timed = !s_board_path.empty();
  const double deadline = timed ? wtime() + rendezvous_timeout() : 0.0;
  const int gen = s_board->barrier_gen.load();
  if (s_board->barrier_count.fetch_add(1) + 1 == s_size) {
    s_board->barrier_count.store(0);
    s_board->barrier_gen.fetch_add(1);
  } else {
    while (s_board->barrier_gen.load() == gen) {
      usleep(50);
      if (timed && wtime() > deadline) rendezvous_failed("barrier");
    }
  }
}

// Wait until rank 0 has published the NCCL id on the board (the role of mpi_bcast), with the same deadline.
void wait_for_nccl_id() {
  if (!s_board) return;
  const bool timed = !s_board_path.empty();
  const double deadline = timed ? wtime() + rendezvous_timeout() : 0.0;
  while (!s_board->id_ready.load()) {
    usleep(100);
    if (timed && wtime() > deadline) rendezvous_failed("NCCL id");
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
