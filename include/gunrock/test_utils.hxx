// test_utils.hxx -- what the test drivers need: --key=value command line parsing,
// device array printing, a wall-clock timer and the two validators.  Same names and call
// signatures as gunrock/tests/test_utils.hxx:17-213.
#pragma once
#include <chrono>
#include <cmath>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "mgpu_compat.hxx"

using namespace std;

namespace gunrock {

// Collects every "--name" / "--name=value" argument; anything else is ignored.
class CommandLineArgs {
    std::vector<std::string> raw;

   protected:
    std::map<std::string, std::string> pairs;

   public:
    CommandLineArgs(int _argc, char **_argv) : raw(_argv, _argv + _argc) {
        for (size_t i = 1; i < raw.size(); ++i) {
            const std::string &arg = raw[i];
            if (arg.size() < 2 || arg.compare(0, 2, "--") != 0) continue;
            const size_t eq = arg.find('=');
            if (eq == std::string::npos) pairs[arg.substr(2)] = "";
            else pairs[arg.substr(2, eq - 2)] = arg.substr(eq + 1);
        }
    }

    bool CheckCmdLineFlag(const char *arg_name) { return pairs.count(arg_name) != 0; }

    // --name=value  ->  val (left untouched when the flag is absent)
    template <typename T>
    void GetCmdLineArgument(const char *arg_name, T &val) {
        auto it = pairs.find(arg_name);
        if (it == pairs.end()) return;
        std::istringstream in(it->second);
        in >> val;
    }

    // --name=v0,v1,...  ->  vals (replaces any defaults when the flag is present)
    template <typename T>
    void GetCmdLineArguments(const char *arg_name, std::vector<T> &vals) {
        auto it = pairs.find(arg_name);
        if (it == pairs.end()) return;
        vals.clear();
        std::istringstream in(it->second);
        std::string piece;
        while (std::getline(in, piece, ',')) {
            if (piece.empty()) continue;
            std::istringstream one(piece);
            T v;
            one >> v;
            vals.push_back(v);
        }
    }

    int ParsedArgc() { return (int)pairs.size(); }

    std::string GetEntireCommandLine() const {
        std::string all;
        for (size_t i = 0; i < raw.size(); ++i) all += (i ? " " : "") + raw[i];
        return all;
    }

    template <typename T>
    void ParseArgument(const char *name, T &val) {
        if (CheckCmdLineFlag(name)) GetCmdLineArgument(name, val);
    }
};

template <typename type_t>
cudaError_t display_device_data(const type_t *data, std::size_t length) {
    std::vector<type_t> host;
    const cudaError_t rc = mgpu::dtoh(host, data, length);
    if (rc != cudaSuccess) return rc;
    for (const type_t &x : host) std::cout << x << ' ';
    std::cout << std::endl;
    return cudaSuccess;
}

// Host wall clock around enact(), as the reference times it (test_bfs.cu:38-42).
class test_timer_t {
    std::chrono::system_clock::time_point t0;
    bool running = false;

   public:
    void start() {
        if (running) return;
        running = true;
        t0 = std::chrono::system_clock::now();
    }
    double end() {
        if (!running) return 0.0;
        return std::chrono::duration<double>(std::chrono::system_clock::now() - t0).count();
    }
};

// exact comparison (labels, predecessors)
inline bool validate(std::vector<int> &gpu_vals, std::vector<int> &cpu_vals) {
    return gpu_vals.size() == cpu_vals.size() && std::equal(gpu_vals.begin(), gpu_vals.end(), cpu_vals.begin());
}

// absolute tolerance 0.01 (test_utils.hxx:204-213)
inline bool validate(std::vector<float> &gpu_vals, std::vector<float> &cpu_vals) {
    if (gpu_vals.size() != cpu_vals.size()) return false;
    for (size_t i = 0; i < gpu_vals.size(); ++i)
        if (!(std::fabs(gpu_vals[i] - cpu_vals[i]) < 0.01f)) return false;
    return true;
}

}  // namespace gunrock
