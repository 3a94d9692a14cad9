// hsk_mpirun — starts N processes of one node as the ranks of a job for the bundled MPI stand-in
// (hysortk_b200/shim/mpi.h): the role `mpirun -np N` plays for the reference (reference README.md:36-52).
//
//   hsk_mpirun -n N [-t THREADS_PER_RANK] [-g] command [args...]
//
// Every child gets HSK_MPI_SIZE / HSK_MPI_RANK / HSK_MPI_SESSION (a name unique to this launch), LOCAL_RANK (the GPU the
// rank uses, hysortk.cpp: local_device_for) and, with -t, OMP_NUM_THREADS; -g also sets CUDA_VISIBLE_DEVICES=<rank>.
// The exit status is the first non-zero child status; when a child fails the others are terminated, and the
// shared-memory object is removed in every case.
#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char **argv)
{
    int n = 1, threads = 0, a = 1;
    bool pin_gpu = false;
    for (; a < argc && argv[a][0] == '-'; ++a) {
        if (!strcmp(argv[a], "-n") && a + 1 < argc) n = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-t") && a + 1 < argc) threads = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-g")) pin_gpu = true;
        else { fprintf(stderr, "usage: hsk_mpirun -n N [-t THREADS] [-g] command [args...]\n"); return 2; }
    }
    if (a >= argc || n < 1) { fprintf(stderr, "usage: hsk_mpirun -n N [-t THREADS] [-g] command [args...]\n"); return 2; }
    const long long stamp = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    const std::string session = std::to_string((long long)getpid()) + "_" + std::to_string(stamp);
    std::vector<pid_t> kids;
    for (int r = 0; r < n; ++r) {
        const pid_t pid = fork();
        if (pid < 0) { perror("fork"); break; }
        if (pid == 0) {
            setenv("HSK_MPI_SIZE", std::to_string(n).c_str(), 1);
            setenv("HSK_MPI_RANK", std::to_string(r).c_str(), 1);
            setenv("HSK_MPI_SESSION", session.c_str(), 1);
            setenv("LOCAL_RANK", std::to_string(r).c_str(), 1);
            if (threads > 0) setenv("OMP_NUM_THREADS", std::to_string(threads).c_str(), 1);
            if (pin_gpu) { setenv("CUDA_VISIBLE_DEVICES", std::to_string(r).c_str(), 1); setenv("LOCAL_RANK", "0", 1); }
            execvp(argv[a], argv + a);
            perror("execvp");
            _exit(127);
        }
        kids.push_back(pid);
    }
    int rc = (int)kids.size() == n ? 0 : 1;
    size_t left = kids.size();
    while (left) {
        int st = 0;
        const pid_t pid = wait(&st);
        if (pid < 0) break;
        --left;
        const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
        for (auto &k : kids) if (k == pid) k = -1;
        if (code != 0 && rc == 0) {
            rc = code;
            for (pid_t k : kids) if (k > 0) kill(k, SIGTERM);   // our own children only
        }
    }
    shm_unlink(("/hsk_mpi_" + session).c_str());
    return rc;
}
