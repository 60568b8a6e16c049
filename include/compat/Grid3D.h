/* include/compat/Grid3D.h -- shim with the public interface of the reference's Grid3D
   (NOCUDA_TESI/POISSON_3D(TESI)/Grid3D.h:4-38).  Same include guard as the reference header. */
#ifndef GRID3D_H
#define GRID3D_H

#include "mg_compat_common.h"

class Grid3D
{
	public:
		float* h_v; // approximate solution (host mirror of the device field)
		float* h_f; // right-hand side

		int sizeX;
		int sizeY;
		int sizeZ;
		int* sizeXYZ;

		float h_x;
		float h_y;
		float h_z;

		float x_a;
		float x_b;
		float y_a;
		float y_b;
		float z_a;
		float z_b;

#ifdef MG_COMPAT_CUDA_TESI
		/* CUDA_TESI face (CUDA_TESI/CUDA Poisson 3D/Grid3D.h:11,26-27): the size triplet and the two fields as DEVICE arrays in
		   the reference's dense layout.  They are mirrors of the engine's (colour-split) level, refreshed by pull() after every
		   MultiGrid3D call and read back by push() before it -- a caller may hand them to the operators or modify them. */
		int* d_sizeXYZ;
		float* d_v;
		float* d_f;
		void PrintDiffApproxReal(int diff_fd) { PrintDiff(diff_fd); } /* C3/Grid3D.h:34 */
#endif

		/* standalone construction, as in the reference: the fields are initialised by the engine
		   (InitV / InitF run on the device) */
		Grid3D(int sizeXYZ_[], float range[]) { setup(sizeXYZ_, range); InitV(); InitF(); upload(); }
		/* used by MultiGrid3D: wraps level `level` of an existing engine handle */
		Grid3D(int sizeXYZ_[], float range[], mg3d_t* mg, int level) { setup(sizeXYZ_, range); pull(mg, level); }
#ifndef MG_COMPAT_CUDA_TESI
		/* ... of a NON-CUBIC hierarchy (sizeX != sizeY != sizeZ: the reference asserts these away, N3/Grid3D.cpp:10-11; here they
		   run through the mg3b_* entry points) */
		Grid3D(int sizeXYZ_[], float range[], mg3b_t* mg, int level) { setup(sizeXYZ_, range); pull(mg, level); }
		void pull(mg3b_t* mg, int level)
		{
			MG_CHECK(mg3b_get_field(mg, level, MG_FIELD_V, h_v));
			MG_CHECK(mg3b_get_field(mg, level, MG_FIELD_F, h_f));
		}
		void push(mg3b_t* mg, int level) const
		{
			MG_CHECK(mg3b_set_field(mg, level, MG_FIELD_V, h_v));
			MG_CHECK(mg3b_set_field(mg, level, MG_FIELD_F, h_f));
		}
#endif
		~Grid3D()
		{
			free(h_v); free(h_f); free(sizeXYZ);
#ifdef MG_COMPAT_CUDA_TESI
			cudaFree(d_v); cudaFree(d_f); cudaFree(d_sizeXYZ);
#endif
		}

		void InitV() { fetch(MG_FIELD_V); }
		void InitF() { fetch(MG_FIELD_F); }

		void pull(mg3d_t* mg, int level)
		{
#ifdef MG_COMPAT_CUDA_TESI
			MG_CHECK(mg3d_get_field_device(mg, level, MG_FIELD_V, d_v));
			MG_CHECK(mg3d_get_field_device(mg, level, MG_FIELD_F, d_f));
			MG_CUDA_CHECK(cudaMemcpy(h_v, d_v, bytes(), cudaMemcpyDeviceToHost)); /* host mirrors: only the Print* dumps read them */
			MG_CUDA_CHECK(cudaMemcpy(h_f, d_f, bytes(), cudaMemcpyDeviceToHost));
#else
			MG_CHECK(mg3d_get_field(mg, level, MG_FIELD_V, h_v));
			MG_CHECK(mg3d_get_field(mg, level, MG_FIELD_F, h_f));
#endif
		}
		void push(mg3d_t* mg, int level) const
		{
#ifdef MG_COMPAT_CUDA_TESI
			MG_CHECK(mg3d_set_field_device(mg, level, MG_FIELD_V, d_v));
			MG_CHECK(mg3d_set_field_device(mg, level, MG_FIELD_F, d_f));
#else
			MG_CHECK(mg3d_set_field(mg, level, MG_FIELD_V, h_v));
			MG_CHECK(mg3d_set_field(mg, level, MG_FIELD_F, h_f));
#endif
		}

		void PrintGrid_v(int logfd) { dump(logfd, h_v, "value"); }
		void PrintGrid_f(int logfd) { dump(logfd, h_f, "value"); }
		void PrintDiff(int logfd)
		{
			const double pi = 3.141592653589793;
			char line[200];
			for (int y = 0; y < sizeY; y++)
				for (int x = 0; x < sizeX; x++)
					for (int z = 0; z < sizeZ; z++) {
						float px = x_a + x * h_x, py = y_a + y * h_y, pz = z_a + z * h_z;
						float exact = sin(pi * px) * sin(pi * py) * sin(pi * pz);
						float diff = exact - h_v[x + y * sizeX + z * sizeX * sizeY];
						snprintf(line, sizeof line, "posY: %d posX: %d posZ: %d diff: %f\n", y, x, z, diff);
						mg_compat_write(logfd, line);
					}
		}
		double MaxAbsError() const
		{
			const double pi = 3.141592653589793;
			double m = 0;
			for (int z = 0; z < sizeZ; z++)
				for (int y = 0; y < sizeY; y++)
					for (int x = 0; x < sizeX; x++) {
						float px = x_a + x * h_x, py = y_a + y * h_y, pz = z_a + z * h_z;
						double d = fabs(sin(pi * px) * sin(pi * py) * sin(pi * pz) - (double)h_v[x + y * sizeX + z * sizeX * sizeY]);
						if (d > m) m = d;
					}
			return m;
		}

	private:
		float range_[6];
		void setup(int s[], float range[])
		{
			sizeX = s[0]; sizeY = s[1]; sizeZ = s[2];
			sizeXYZ = (int*)malloc(3 * sizeof(int));
			sizeXYZ[0] = sizeX; sizeXYZ[1] = sizeY; sizeXYZ[2] = sizeZ;
			for (int i = 0; i < 6; i++) range_[i] = range[i];
			x_a = range[0]; x_b = range[1]; y_a = range[2]; y_b = range[3]; z_a = range[4]; z_b = range[5];
			h_x = (x_b - x_a) / (float)(sizeX - 1);
			h_y = (y_b - y_a) / (float)(sizeY - 1);
			h_z = (z_b - z_a) / (float)(sizeZ - 1);
			size_t tot = (size_t)sizeX * sizeY * sizeZ;
			h_v = (float*)malloc(tot * sizeof(float));
			h_f = (float*)malloc(tot * sizeof(float));
#ifdef MG_COMPAT_CUDA_TESI
			MG_CUDA_CHECK(cudaMalloc((void**)&d_v, tot * sizeof(float)));
			MG_CUDA_CHECK(cudaMalloc((void**)&d_f, tot * sizeof(float)));
			MG_CUDA_CHECK(cudaMalloc((void**)&d_sizeXYZ, 3 * sizeof(int)));
			MG_CUDA_CHECK(cudaMemcpy(d_sizeXYZ, sizeXYZ, 3 * sizeof(int), cudaMemcpyHostToDevice));
#endif
		}
		size_t bytes() const { return (size_t)sizeX * sizeY * sizeZ * sizeof(float); }
		void upload()
		{
#ifdef MG_COMPAT_CUDA_TESI
			MG_CUDA_CHECK(cudaMemcpy(d_v, h_v, bytes(), cudaMemcpyHostToDevice));
			MG_CUDA_CHECK(cudaMemcpy(d_f, h_f, bytes(), cudaMemcpyHostToDevice));
#endif
		}
		void fetch(int field)
		{
			double r[6];
			for (int i = 0; i < 6; i++) r[i] = range_[i];
			if (sizeX != sizeY || sizeX != sizeZ) { /* non-cubic: mg3b_* */
				mg3b_t* mb = 0;
				MG_CHECK(mg3b_create(&mb, sizeXYZ, r, MG_F32, MG_REF_COMPAT));
				MG_CHECK(mg3b_get_field(mb, 0, field, field == MG_FIELD_V ? h_v : h_f));
				mg3b_destroy(mb);
				return;
			}
			mg3d_t* mg = 0;
			MG_CHECK(mg3d_create(&mg, sizeXYZ, r, MG_F32, MG_REF_COMPAT));
			MG_CHECK(mg3d_get_field(mg, 0, field, field == MG_FIELD_V ? h_v : h_f));
			mg3d_destroy(mg);
		}
		void dump(int logfd, const float* a, const char* what)
		{
			char line[200];
			for (int y = 0; y < sizeY; y++)
				for (int x = 0; x < sizeX; x++)
					for (int z = 0; z < sizeZ; z++) {
						snprintf(line, sizeof line, "posY: %d posX: %d posZ: %d %s: %f\n", y, x, z, what, a[x + y * sizeX + z * sizeX * sizeY]);
						mg_compat_write(logfd, line);
					}
		}
};
#endif
