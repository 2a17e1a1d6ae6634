package de.jstacs.gpu;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.foreign.ValueLayout;
import java.lang.invoke.MethodHandle;

/**
 * Panama FFM (JDK &gt;= 22) binding of <code>libjstacs_hmm_b200.so</code> (C ABI: <code>include/jstacs_hmm.h</code>).
 * One downcall handle per exported function; no JNI glue is needed because the ABI only uses pointers, 32/64 bit
 * integers and doubles.  The library is located through the system property <code>jstacs.hmm.lib</code> or
 * <code>java.library.path</code>, the same convention as <code>System.loadLibrary("UserTime")</code> in
 * <code>de/jstacs/utils/UserTime.java:47</code> — but, unlike the timer, there is NO fallback: a missing library or
 * device is an error (north star: no CPU fallback).
 */
public final class JstacsHmmNative {

	public static final int JH_OK = 0, JH_ERR_INVALID = -1, JH_ERR_UNSUPPORTED = -2, JH_ERR_CUDA = -3, JH_ERR_ALPHABET = -4,
			JH_ERR_NOMEM = -5, JH_ERR_IMPOSSIBLE = -6, JH_ERR_COMM = -7;
	public static final int JH_MODE_VITERBI = 0, JH_MODE_BAUM_WELCH = 1, JH_TERM_ITERATIONS = 0, JH_TERM_SMALL_DIFFERENCE = 1;
	public static final int JH_ABI_VERSION = 2;

	private static final Linker LINKER = Linker.nativeLinker();
	private static final SymbolLookup LIB;
	static {
		String path = System.getProperty( "jstacs.hmm.lib" );
		if( path != null ) {
			LIB = SymbolLookup.libraryLookup( path, Arena.global() );
		} else {
			System.loadLibrary( "jstacs_hmm_b200" );
			LIB = SymbolLookup.loaderLookup();
		}
	}

	private static final ValueLayout.OfInt I = ValueLayout.JAVA_INT;
	private static final ValueLayout.OfLong J = ValueLayout.JAVA_LONG;
	private static final java.lang.foreign.AddressLayout P = ValueLayout.ADDRESS;

	private static MethodHandle h( String name, FunctionDescriptor fd ) {
		return LINKER.downcallHandle( LIB.find( name ).orElseThrow( () -> new UnsatisfiedLinkError( name ) ), fd );
	}

	public static final MethodHandle jh_abi_version = h( "jh_abi_version", FunctionDescriptor.of( I ) );
	public static final MethodHandle jh_last_error = h( "jh_last_error", FunctionDescriptor.of( P ) );
	public static final MethodHandle jh_device_count = h( "jh_device_count", FunctionDescriptor.of( I ) );
	public static final MethodHandle jh_set_device = h( "jh_set_device", FunctionDescriptor.of( I, I ) );
	public static final MethodHandle jh_model_create = h( "jh_model_create", FunctionDescriptor.of( I, P, P ) );
	public static final MethodHandle jh_model_destroy = h( "jh_model_destroy", FunctionDescriptor.ofVoid( P ) );
	public static final MethodHandle jh_model_set_params = h( "jh_model_set_params", FunctionDescriptor.of( I, P, P, P, P, P ) );
	public static final MethodHandle jh_model_get_params = h( "jh_model_get_params", FunctionDescriptor.of( I, P, P, P, P, P ) );
	public static final MethodHandle jh_model_sizes = h( "jh_model_sizes", FunctionDescriptor.of( I, P, P ) );
	public static final MethodHandle jh_dataset_create = h( "jh_dataset_create", FunctionDescriptor.of( I, P, P, J, P, I, P ) );
	public static final MethodHandle jh_dataset_destroy = h( "jh_dataset_destroy", FunctionDescriptor.ofVoid( P ) );
	public static final MethodHandle jh_dataset_set_weights = h( "jh_dataset_set_weights", FunctionDescriptor.of( I, P, P ) );
	public static final MethodHandle jh_loglik = h( "jh_loglik", FunctionDescriptor.of( I, P, P, P ) );
	public static final MethodHandle jh_loglik_windows = h( "jh_loglik_windows", FunctionDescriptor.of( I, P, P, J, P, P, P, P ) );
	public static final MethodHandle jh_viterbi = h( "jh_viterbi", FunctionDescriptor.of( I, P, P, P, P ) );
	public static final MethodHandle jh_viterbi_u8 = h( "jh_viterbi_u8", FunctionDescriptor.of( I, P, P, P, P ) );
	public static final MethodHandle jh_viterbi_paths = h( "jh_viterbi_paths", FunctionDescriptor.of( I, P, P, P, P, P, P ) );
	public static final MethodHandle jh_viterbi_windows = h( "jh_viterbi_windows", FunctionDescriptor.of( I, P, P, J, P, P, P, P, P, P, P ) );
	public static final MethodHandle jh_posterior_window = h( "jh_posterior_window", FunctionDescriptor.of( I, P, P, J, J, J, P ) );
	public static final MethodHandle jh_posterior = h( "jh_posterior", FunctionDescriptor.of( I, P, P, J, P ) );
	public static final MethodHandle jh_posterior_decode = h( "jh_posterior_decode", FunctionDescriptor.of( I, P, P, P, P ) );
	public static final MethodHandle jh_posterior_decode_u8 = h( "jh_posterior_decode_u8", FunctionDescriptor.of( I, P, P, P, P ) );
	public static final MethodHandle jh_estep = h( "jh_estep", FunctionDescriptor.of( I, P, P, I, P, P, P ) );
	public static final MethodHandle jh_mstep = h( "jh_mstep", FunctionDescriptor.of( I, P ) );
	public static final MethodHandle jh_baum_welch = h( "jh_baum_welch", FunctionDescriptor.of( I, P, P, P, P, I, P ) );
	public static final MethodHandle jh_last_objective = h( "jh_last_objective", FunctionDescriptor.of( I, P, P ) );
	public static final MethodHandle jh_model_log_prior = h( "jh_model_log_prior", FunctionDescriptor.of( I, P, P ) );
	public static final MethodHandle jh_model_set_option = h( "jh_model_set_option", FunctionDescriptor.of( I, P, P, J ) );
	public static final MethodHandle jh_model_set_stream = h( "jh_model_set_stream", FunctionDescriptor.of( I, P, P ) );
	public static final MethodHandle jh_set_option = h( "jh_set_option", FunctionDescriptor.of( I, P, J ) );
	public static final MethodHandle jh_dataset_create_packed2 = h( "jh_dataset_create_packed2", FunctionDescriptor.of( I, P, P, P, J, P, I, P ) );
	public static final MethodHandle jh_model_set_states_groups = h( "jh_model_set_states_groups", FunctionDescriptor.of( I, P, I, P, P ) );
	public static final MethodHandle jh_dataset_set_groups = h( "jh_dataset_set_groups", FunctionDescriptor.of( I, P, P ) );
	public static final MethodHandle jh_dataset_info = h( "jh_dataset_info", FunctionDescriptor.of( I, P, P, P, P ) );

	static {
		try {
			int v = (int) jh_abi_version.invokeExact();
			if( v != JH_ABI_VERSION ) {
				throw new UnsatisfiedLinkError( "libjstacs_hmm_b200 has ABI version " + v + ", this binding needs " + JH_ABI_VERSION );
			}
		} catch( Error e ) {
			throw e;
		} catch( Throwable t ) {
			throw new UnsatisfiedLinkError( t.toString() );
		}
	}

	/** Turns a jh_status into the exception type the reference throws at the same place (SURVEY.md §8b "Errors"). */
	public static void check( int status ) throws Exception {
		if( status == JH_OK ) {
			return;
		}
		String msg;
		try {
			MemorySegment s = (MemorySegment) jh_last_error.invokeExact();
			msg = s.reinterpret( 1024 ).getString( 0 );
		} catch( Throwable t ) {
			msg = "native error " + status;
		}
		switch( status ) {
			case JH_ERR_ALPHABET:
				throw new de.jstacs.data.WrongAlphabetException( msg ); // AbstractHMM.java:991-993
			case JH_ERR_INVALID:
			case JH_ERR_IMPOSSIBLE:
				throw new IllegalArgumentException( msg ); // e.g. "Impossible path", HigherOrderHMM.java:659
			case JH_ERR_UNSUPPORTED:
				throw new UnsupportedOperationException( msg );
			case JH_ERR_NOMEM:
				throw new OutOfMemoryError( msg );
			default:
				throw new RuntimeException( msg ); // HigherOrderHMM.java:1197-1227 wraps worker failures the same way
		}
	}

	private JstacsHmmNative() {}
}
