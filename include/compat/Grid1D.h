/* include/compat/Grid1D.h -- shim with the public interface of the reference's Grid1D
   (NOCUDA_TESI/EQUAZIONE 1D/Grid1D.h:4-26).  Same include guard as the reference header. */
#ifndef GRID1D_H
#define GRID1D_H

#include "mg_compat_common.h"

class Grid1D
{
	public:
		float* h_v;
		float* h_f;

		int sizeX;
		float h_x;

		float x_a;
		float x_b;

#ifdef MG_COMPAT_CUDA_TESI
		/* CUDA_TESI face (CUDA_TESI/CUDA 1D/Grid1D.h:15-18): the two fields as DEVICE arrays -- the engine's own storage of the
		   level.  NULL for a Grid1D constructed on its own. */
		float* d_v;
		float* d_f;
#endif

		Grid1D(int sizeX_, float range[]) { setup(sizeX_, range); InitV(); InitF(); }
		Grid1D(int sizeX_, float range[], mg1d_t* mg, int level) { setup(sizeX_, range); attach(mg, level); pull(mg, level); }
		~Grid1D() { free(h_v); free(h_f); }

		void InitV() { fetch(MG_FIELD_V); }
		void InitF() { fetch(MG_FIELD_F); }

		void pull(mg1d_t* mg, int level)
		{
			MG_CHECK(mg1d_get_field(mg, level, MG_FIELD_V, h_v));
			MG_CHECK(mg1d_get_field(mg, level, MG_FIELD_F, h_f));
		}
		void push(mg1d_t* mg, int level) const
		{
#ifndef MG_COMPAT_CUDA_TESI /* (CUDA_TESI face: the device arrays are the engine's own, the host arrays only feed the dumps) */
			MG_CHECK(mg1d_set_field(mg, level, MG_FIELD_V, h_v));
			MG_CHECK(mg1d_set_field(mg, level, MG_FIELD_F, h_f));
#else
			(void)mg; (void)level;
#endif
		}
		void attach(mg1d_t* mg, int level)
		{
#ifdef MG_COMPAT_CUDA_TESI
			void *pv = 0, *pf = 0;
			MG_CHECK(mg1d_level_device_ptr(mg, level, MG_FIELD_V, &pv));
			MG_CHECK(mg1d_level_device_ptr(mg, level, MG_FIELD_F, &pf));
			d_v = (float*)pv; d_f = (float*)pf;
#else
			(void)mg; (void)level;
#endif
		}

		void PrintDiffApproxReal(int diff_fd)
		{
			char line[100];
			for (int x = 0; x < sizeX; x++) {
				float xj = x_a + x * h_x;
				float exact = (expf(xj) + xj - 3) / (1 + expf(-xj));
				snprintf(line, sizeof line, "xj: %f diff: %f\n", xj, h_v[x] - exact);
				mg_compat_write(diff_fd, line);
			}
		}
		double MaxAbsError() const
		{
			double m = 0;
			for (int x = 0; x < sizeX; x++) {
				double xj = x_a + x * (double)h_x;
				double d = fabs(h_v[x] - (exp(xj) + xj - 3) / (1 + exp(-xj)));
				if (d > m) m = d;
			}
			return m;
		}
		void PrintGrid_v(int logfd) { dump(logfd, h_v); }
		void PrintGrid_f(int logfd) { dump(logfd, h_f); }

	private:
		float range_[2];
		void setup(int n, float range[])
		{
			sizeX = n;
			range_[0] = range[0]; range_[1] = range[1];
			x_a = range[0]; x_b = range[1];
			h_x = (x_b - x_a) / (float)(sizeX - 1);
			h_v = (float*)malloc((size_t)n * sizeof(float));
			h_f = (float*)malloc((size_t)n * sizeof(float));
#ifdef MG_COMPAT_CUDA_TESI
			d_v = d_f = 0;
#endif
		}
		void fetch(int field)
		{
			double r[2] = {range_[0], range_[1]};
			mg1d_t* mg = 0;
			MG_CHECK(mg1d_create(&mg, sizeX, r, MG_F32, MG_REF_COMPAT));
			MG_CHECK(mg1d_get_field(mg, 0, field, field == MG_FIELD_V ? h_v : h_f));
			mg1d_destroy(mg);
		}
		void dump(int logfd, const float* a)
		{
			char line[100];
			for (int x = 0; x < sizeX; x++) {
				snprintf(line, sizeof line, "posX: %d value: %f\n", x, a[x]);
				mg_compat_write(logfd, line);
			}
		}
};
#endif
