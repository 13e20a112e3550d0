// field2d.hpp — Array2D / Field2D host mirrors (drop-in for reference src/Array.hpp:7-81 and src/Field2D.hpp).
// Pure host containers used for IO and diagnostics; the hot path keeps its own copies on the device and
// Fields::download()/upload() move data between the two.  Storage is one contiguous row-major block with a
// row-pointer table so that `a[i][j]` and `a[0]` (flat) both work as in the reference.
#pragma once
#include <cmath>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <vector>

template <class T>
class Array2D
{
  public:
    int jmax = 0, lmax = 0;
    T** data = nullptr;

    Array2D() = default;
    Array2D(int rows, int cols) { allocate(rows, cols); }
    Array2D(const Array2D& o) { allocate(o.jmax, o.lmax); store = o.store; }
    Array2D& operator=(const Array2D& o)
    {
        if (this != &o) { allocate(o.jmax, o.lmax); store = o.store; }
        return *this;
    }
    void resize(int rows, int cols) { allocate(rows, cols); }
    T*& operator[](int i) const { return data[i]; }

    void add(const Array2D& o)
    {
        same_shape(o, "Array2D.add(): matrix sizes don't match\n");
        for (size_t k = 0; k < store.size(); k++) store[k] += o.store[k];
    }
    void assign(const Array2D& o)
    {
        same_shape(o, "Array2D.assign(): matrix sizes don't match\n");
        store = o.store;
    }
    void multiply(double f) { for (T& v : store) v *= f; }
    void reset() { for (T& v : store) v = 0; }

  private:
    std::vector<T> store;
    std::vector<T*> rows_;
    void allocate(int rows, int cols)
    {
        jmax = rows;
        lmax = cols;
        store.assign((size_t)rows * cols, T());
        rows_.resize(rows);
        for (int i = 0; i < rows; i++) rows_[i] = store.data() + (size_t)i * cols;
        data = rows_.empty() ? nullptr : rows_.data();
    }
    void same_shape(const Array2D& o, const char* msg) const
    {
        if (jmax != o.jmax || lmax != o.lmax) throw std::runtime_error(msg);
    }
};

class Field2D : public Array2D<double>
{
  public:
    Field2D() = default;
    Field2D(int x_sampl, int z_sampl, double dx_, double dy_, double xmin_ = 0, double ymin_ = 0)
        : Array2D<double>(x_sampl, z_sampl), dx(dx_), dy(dy_), idx(1. / dx_), idy(1. / dy_), xmin(xmin_), ymin(ymin_) {}
    void resize(int x_sampl, int z_sampl, double dx_, double dy_, double xmin_ = 0, double ymin_ = 0)
    {
        dx = dx_; dy = dy_; idx = 1.0 / dx; idy = 1.0 / dy; xmin = xmin_; ymin = ymin_;
        Array2D<double>::resize(x_sampl, z_sampl);
    }
    void resize(Field2D& f) { resize(f.jmax, f.lmax, f.dx, f.dy, f.xmin, f.ymin); }
    double x(int i) { return i * dx + xmin; }
    double y(int j) { return j * dy + ymin; }
    double GetDx() { return dx; }
    double GetDy() { return dy; }
    double GetXMin() { return xmin; }
    double GetYMin() { return ymin; }
    double GetXMax() { return xmin + (jmax - 1) * dx; }
    double GetYMax() { return ymin + (lmax - 1) * dy; }

    // bilinear cloud-in-cell scatter of one charge (Field2D.hpp:45-62), host diagnostics only
    void accumulate(double charge, double px, double py)
    {
        int i, j;
        double u, v;
        locate(px, py, i, j, u, v, "Field2D::accumulate() outside of range\n");
        data[i][j] += (1 - u) * (1 - v) * charge;
        data[i + 1][j] += u * (1 - v) * charge;
        data[i][j + 1] += (1 - u) * v * charge;
        data[i + 1][j + 1] += u * v * charge;
    }
    double interpolate(double px, double py) const
    {
        int i, j;
        double u, v;
        locate(px, py, i, j, u, v, "Field2D::interpolate() outside of range\n");
        return (1 - u) * (1 - v) * data[i][j] + u * (1 - v) * data[i + 1][j] + (1 - u) * v * data[i][j + 1] + u * v * data[i + 1][j + 1];
    }
    // gradient by bilinear interpolation of edge-centred differences (Field2D.hpp:80-168); the device
    // version is gather_E in csrc/push.cu
    void grad(double px, double py, double& gx, double& gy) const
    {
        const double X = (px - xmin) * idx, Y = (py - ymin) * idy;
        {
            const int i = (int)(X + 0.5), j = std::min((int)Y, lmax - 2);
            const double fy = Y - j;
            auto edge = [&](int a, int b) { return (data[a][b] - data[a - 1][b]) * idx; };
            if (i > 0 && i < jmax - 1)
            {
                const double fx = X - i + .5;
                gx = edge(i, j) * (1 - fx) * (1 - fy) + edge(i, j + 1) * (1 - fx) * fy + edge(i + 1, j + 1) * fx * fy + edge(i + 1, j) * fx * (1 - fy);
            }
            else if (i == jmax - 1) gx = edge(i, j) * (1 - fy) + edge(i, j + 1) * fy;
            else if (i == 0) gx = edge(1, j + 1) * fy + edge(1, j) * (1 - fy);
        }
        {
            const int i = std::min((int)X, jmax - 2), j = (int)(Y + 0.5);
            const double fx = X - i;
            auto edge = [&](int a, int b) { return (data[a][b] - data[a][b - 1]) * idy; };
            if (j > 0 && j < lmax - 1)
            {
                const double fy = Y - j + 0.5;
                gy = edge(i, j) * (1 - fx) * (1 - fy) + edge(i + 1, j) * (1 - fy) * fx + edge(i + 1, j + 1) * fx * fy + edge(i, j + 1) * fy * (1 - fx);
            }
            else if (j == lmax - 1) gy = edge(i, j) * (1 - fx) + edge(i + 1, j) * fx;
            else if (j == 0) gy = edge(i + 1, 1) * fx + edge(i, 1) * (1 - fx);
        }
    }
    bool hasnan()
    {
        for (int i = 0; i < jmax; i++)
            for (int j = 0; j < lmax; j++)
                if (std::isnan(data[i][j])) return true;
        return false;
    }
    // "x <tab> y <tab> value" rows, blank line after each x (gnuplot grid format, Field2D.cpp:29-38)
    void print(std::ostream& out = std::cout, double factor = 1.0)
    {
        for (int i = 0; i < jmax; i++)
        {
            for (int j = 0; j < lmax; j++) out << i * dx + xmin << "\t" << j * dy + ymin << "\t" << data[i][j] * factor << std::endl;
            out << std::endl;
        }
    }
    void print(const char* filename, double factor = 1.0)
    {
        std::ofstream out(filename);
        print(out, factor);
    }
    // inverse of print(): three-column file on a regular grid, any row order (Field2D.cpp:46-130)
    void load(const char* filename)
    {
        std::ifstream in(filename);
        if (in.fail()) throw std::runtime_error("Field2D::load(): failed opening file\n");
        std::vector<double> xs, ys, fs;
        for (std::string line; std::getline(in, line);)
        {
            std::istringstream row(line);
            double a, b, c;
            if (row >> a >> b >> c) { xs.push_back(a); ys.push_back(b); fs.push_back(c); }
        }
        if (xs.empty()) throw std::runtime_error("Field2D::load(): failed opening file\n");
        auto axis = [&](const std::vector<double>& v, double& lo, double& step) {
            lo = v.front();
            double hi = v.back();
            size_t k = 1;
            while (k < v.size() && v[k] - v[k - 1] == 0.0) k++;
            step = k < v.size() ? v[k] - v[k - 1] : 1.0;
            if (step < 0) { step = -step; std::swap(lo, hi); }
            return (unsigned)std::lround((hi - lo) / step + 1);
        };
        double x0, y0, sx, sy;
        const unsigned nx = axis(xs, x0, sx), ny = axis(ys, y0, sy);
        if ((size_t)nx * ny != xs.size()) throw std::runtime_error("Fields::load_magnetic_field() wrong size of input vector");
        resize(nx, ny, sx, sy, x0, y0);
        for (int i = 0; i < jmax; i++)
            for (int j = 0; j < lmax; j++) data[i][j] = std::numeric_limits<double>::quiet_NaN();
        for (size_t k = 0; k < xs.size(); k++) data[std::lround((xs[k] - x0) / sx)][std::lround((ys[k] - y0) / sy)] = fs[k];
        if (hasnan()) throw std::runtime_error("Fields::load_magnetic_field() garbage loaded");
    }

  private:
    double dx = 0, dy = 0, idx = 0, idy = 0, xmin = 0, ymin = 0;
    void locate(double px, double py, int& i, int& j, double& u, double& v, const char* msg) const
    {
        px -= xmin;
        py -= ymin;
        i = (int)(px * idx);
        j = (int)(py * idy);
        u = px * idx - i;
        v = py * idy - j;
        if (i < 0 || i > jmax - 1 || j < 0 || j > lmax - 1) throw std::runtime_error(msg);
    }
};

// Histogram with the reference's binning rule (src/histogram.cpp): strict bounds for the bins, running
// totals over everything that was offered
class Histogram
{
  public:
    Histogram(int n, double lo_, double hi_) : bins(n, 0.0), lo(lo_), hi(hi_) {}
    double& operator[](int i) { return bins[i]; }
    double position(int i) { return lo + (hi - lo) * (i + .5) / bins.size(); }
    int N_hist() { return (int)bins.size(); }
    double Min() { return lo; }
    double Max() { return hi; }
    double N_val() { return n_in; }
    int add(double f, double weight = 1.0)
    {
        int j = -1;
        if (f < hi && f > lo)
        {
            j = (int)((f - lo) * bins.size() / (hi - lo));
            bins[j] += weight;
            n_in += weight;
            sum_in += weight * f;
        }
        n_all += weight;
        sum_all += weight * f;
        return j;
    }
    // merge a histogram computed on the device (mag2d_energy_hist): counts per bin and {n_in, sum_in, n_all, sum_all}
    void add_counts(const double* counts, const double stats[4])
    {
        for (size_t k = 0; k < bins.size(); k++) bins[k] += counts[k];
        n_in += stats[0]; sum_in += stats[1]; n_all += stats[2]; sum_all += stats[3];
    }
    void reset()
    {
        std::fill(bins.begin(), bins.end(), 0.0);
        n_in = sum_in = n_all = sum_all = 0;
    }
    double mean() { return sum_in / n_in; }
    double mean_tot() { return sum_all / n_all; }
    double norm() { return N_val() * (hi - lo) / bins.size(); }
    void print(std::ostream& out = std::cout)
    {
        for (size_t k = 0; k < bins.size(); k++) out << position((int)k) << "\t" << bins[k] / norm() << std::endl;
    }
    void print(const char* fname)
    {
        std::ofstream out(fname);
        print(out);
    }

  private:
    std::vector<double> bins;
    double lo, hi;
    double n_in = 0, sum_in = 0, n_all = 0, sum_all = 0;
};
