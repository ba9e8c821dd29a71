// A CabanaPIC-format deck for BASELINE.json configs[3]: periodic 3-D EM box with a bi-Maxwellian
// electron population (hotter across z than along it) -- the thermal anisotropy drives the Weibel
// instability, magnetic energy grows from noise.  Written against the reference's deck interface
// (src/input/deck.h): a custom Particle_Initializer plus the Input_Deck constructor.
// Grid and particle count scale with CPIC_WEIBEL_N (cells per axis, default 32) and
// CPIC_WEIBEL_PPC (default 16), so the same deck serves tests (small) and multi-GPU runs (128^3).
#include "src/input/deck.h"

#include <cstdlib>
#include <random>

class Weibel_Particles : public Particle_Initializer {
   public:
    using real_ = real_t;
    real_ uth_perp, uth_par;
    Weibel_Particles(real_ perp, real_ par) : uth_perp(perp), uth_par(par) {}
    virtual void init(particle_list_t& particles, size_t nx, size_t ny, size_t nz, size_t ng, real_, size_t nppc, real_ w, real_,
                      real_, real_, real_) {
        auto px = Cabana::slice<PositionX>(particles);
        auto py = Cabana::slice<PositionY>(particles);
        auto pz = Cabana::slice<PositionZ>(particles);
        auto ux = Cabana::slice<VelocityX>(particles);
        auto uy = Cabana::slice<VelocityY>(particles);
        auto uz = Cabana::slice<VelocityZ>(particles);
        auto weight = Cabana::slice<Weight>(particles);
        auto cell = Cabana::slice<Cell_Index>(particles);
        std::mt19937_64 rng(20260101);
        std::uniform_real_distribution<double> uni(-1.0, 1.0);
        std::normal_distribution<double> gauss(0.0, 1.0);
        const size_t n = particles.size();
        for (size_t k = 0; k < n; ++k) {          // serial, in particle order: reproducible
            const size_t c = k / nppc;             // cell by cell: the store starts cell-sorted
            const size_t ix = c % nx, iy = (c / nx) % ny, iz = c / (nx * ny);
            px(k) = uni(rng); py(k) = uni(rng); pz(k) = uni(rng);
            ux(k) = uth_perp * gauss(rng);
            uy(k) = uth_perp * gauss(rng);
            uz(k) = uth_par * gauss(rng);
            weight(k) = w;
            cell(k) = VOXEL(ix + ng, iy + ng, iz + ng, nx, ny, nz, ng);
        }
    }
};

Input_Deck::Input_Deck() {
    const char* e = std::getenv("CPIC_WEIBEL_N");
    const size_t n = e ? std::atoi(e) : 32;
    e = std::getenv("CPIC_WEIBEL_PPC");
    nx = ny = nz = n;
    nppc = e ? std::atoi(e) : 16;
    num_steps = 200;
    const real_ cell = 0.25;                       // d_e per cell
    len_x_global = len_y_global = len_z_global = n * cell;
    dt = 0.99 * courant_length(len_x_global, len_y_global, len_z_global, nx, ny, nz) / c;
    n0 = 1.0;
    v0 = 0.0;
    particle_initer = new Weibel_Particles(0.30, 0.06);    // T_perp / T_par = 25
}
